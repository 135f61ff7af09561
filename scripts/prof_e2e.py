"""Time the host<->device legs of the e2e bench step (C2 shape) separately."""
import ctypes as C, sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
S = pkg.solver
cfg = pkg.problems.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "C2"]
opt = pkg.Optimizer()
for k, v in dict(cfg["options"], verb=0, device=0).items():
    opt.set_attribute(k, v)
opt.copy_to(pkg.raw_from_sdpa_arrays(*cfg["gen"]()))
s, ha = opt.solver, opt.halpha
S.setup_solver(s, ha)
md = s.model
S.initial_point(s)
for _ in range(2):
    S.myIPstep(s, ha)
y, X, xl = S.get_solution(s)
PD = C.POINTER(C.c_double)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
Xh = [np.asfortranarray(pin(x.T).T) for x in X]
Sh = [np.asfortranarray(pin(x.T).T) for x in X]
yh = pin(y)
def t(f, n=3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("set_iterate pinned  %.1f ms" % t(lambda: S.set_iterate(s, Xh, Sh, yh, xl, xl)))
print("get_solution pinned %.1f ms" % t(lambda: S.get_solution(s, out=(yh, Xh, xl))))
print("get_solution fresh  %.1f ms" % t(lambda: S.get_solution(s)))
a = torch.empty(5000 * 5000, dtype=torch.float64).pin_memory(); d = torch.empty_like(a, device="cuda")
print("torch H2D 200MB pinned %.1f ms" % t(lambda: d.copy_(a, non_blocking=True)))
print("torch D2H 200MB pinned %.1f ms" % t(lambda: a.copy_(d, non_blocking=True)))
print("myIPstep %.1f ms" % t(lambda: S.myIPstep(s, ha), 2))
print("check_convergence %.1f ms" % t(lambda: S.check_convergence(s), 2))
