// TEST INFRASTRUCTURE ONLY: the NCCL types csrc/solver.cu names (it resolves the functions with dlsym at run time and never does
// on one GPU); lets the host code compile for the CPU emulator.
#pragma once
#include <stddef.h>
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0, ncclUnhandledCudaError = 1 } ncclResult_t;
typedef enum { ncclInt8 = 0, ncclChar = 0, ncclFloat32 = 7, ncclFloat = 7, ncclFloat64 = 8, ncclDouble = 8 } ncclDataType_t;
typedef enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 } ncclRedOp_t;
