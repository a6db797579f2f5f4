"""TEST INFRASTRUCTURE ONLY: rewrites the `kernel<<<grid, block, smem, stream>>>(args);` launches of a .cu file into calls of the
CPU SIMT emulator (tests/cuda_emu/cuda_emu.h), so that the library's HOST code (plan set-up, kernel selection, chunking, the
ADMM / CG drivers, the C ABI) can be compiled by g++ and logic-checked without a GPU.  The output goes to a temporary
directory of the test that asked for it; nothing in the product reads it.

    kernel<T, Cfg><<<g, b, s, st>>>(a, b);   ->   ::cuda_emu::launch(dim3(g), dim3(b), (size_t)(s), [&] { kernel<T, Cfg>(a, b); });
"""
import sys


def _match_back(s, i, open_c, close_c):
    """s[i] == close_c: index of the matching open_c."""
    depth = 0
    while i >= 0:
        if s[i] == close_c:
            depth += 1
        elif s[i] == open_c:
            depth -= 1
            if depth == 0:
                return i
        i -= 1
    raise ValueError("unbalanced %s%s" % (open_c, close_c))


def _match_fwd(s, i, open_c, close_c):
    """s[i] == open_c: index of the matching close_c."""
    depth = 0
    while i < len(s):
        if s[i] == open_c:
            depth += 1
        elif s[i] == close_c:
            depth -= 1
            if depth == 0:
                return i
        i += 1
    raise ValueError("unbalanced %s%s" % (open_c, close_c))


def _split_top(s):
    out, depth, cur = [], 0, ""
    for i, ch in enumerate(s):
        if ch in "([{<":
            depth += 1
        elif ch in ")]}" or (ch == ">" and s[i - 1:i] != "-"):   # '->' is not a closing bracket
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def translate(src):
    out, pos, count = [], 0, 0
    while True:
        k = src.find("<<<", pos)
        if k < 0:
            out.append(src[pos:])
            break
        # kernel expression: identifier (with ::) and an optional template argument list, just before <<<
        j = k - 1
        if src[j] == ">":
            j = _match_back(src, j, "<", ">") - 1
        while j >= 0 and (src[j].isalnum() or src[j] in "_:"):
            j -= 1
        kern = src[j + 1:k]
        e = src.find(">>>", k)
        cfg = _split_top(src[k + 3:e])
        if not (2 <= len(cfg) <= 4):
            raise ValueError("launch configuration %r" % src[k:e + 3])
        a0 = e + 3
        while src[a0].isspace():
            a0 += 1
        if src[a0] != "(":
            raise ValueError("no argument list after %r" % src[k:e + 3])
        a1 = _match_fwd(src, a0, "(", ")")
        if src[a1 + 1] != ";":
            raise ValueError("launch is not a statement: %r" % src[j + 1:a1 + 2])
        smem = cfg[2] if len(cfg) >= 3 else "0"
        out.append(src[pos:j + 1])
        out.append("::cuda_emu::launch(dim3(%s), dim3(%s), (size_t)(%s), [&] { %s%s; })" % (cfg[0], cfg[1], smem, kern, src[a0:a1 + 1]))
        pos = a1 + 1
        count += 1
    return "".join(out), count


if __name__ == "__main__":
    text, n = translate(open(sys.argv[1]).read())
    open(sys.argv[2], "w").write(text)
    print("%s: %d launches rewritten" % (sys.argv[1], n))
