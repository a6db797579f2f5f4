import sys, numpy as np
sys.path.insert(0, '.')
import multivartv_b200 as mv
rng = np.random.default_rng(1)
for dims in ([256,256],[512,512],[640,512],[1024,1024],[1024,64],[2048,2048],[96,96,96],[40,40,40]):
    N = int(np.prod(dims)); p = len(dims)
    x = rng.random((N, p)); y = np.prod(x > 0.5, axis=1)*1.0 + 0.5*rng.standard_normal(N)
    axes = [np.linspace(0,1,d) for d in dims]
    with mv.Plan(dims) as pl:
        pl.set_points(x, y, axes)
        try:
            a = pl.solve(1.0, mode="rcpp", max_passes=3, precond=mv.PRECOND_JACOBI, cg_maxit=2000)
        except Exception as e:
            print(dims, 'jacobi failed', e); continue
        try:
            b = pl.solve(1.0, mode="rcpp", max_passes=3, precond=mv.PRECOND_CHEB1, cg_maxit=2000)
            print(dims, 'jacobi inner', a['inner_iters'], 'cheb1 inner', b['inner_iters'], 'max diff', np.abs(a['theta']-b['theta']).max())
        except Exception as e:
            print(dims, 'jacobi inner', a['inner_iters'], 'cheb1 FAILED', str(e)[:80])
