"""Dev script: first end-to-end GPU check (correctness vs oracle/reference + coarse timing)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle
from oracle import ref_cufinufft as ref
import jax_finufft_b200 as J
from jax_finufft_b200.plan import Plan

dev = "cuda"
rng = np.random.default_rng(657)
def T(a, dt): return torch.as_tensor(np.ascontiguousarray(a)).to(dt).to(dev)

print("== small cases through the public API vs direct NUDFT")
for x64 in (False, True):
    rd, cd = (torch.float64, torch.complex128) if x64 else (torch.float32, torch.complex64)
    eps = 1e-10 if x64 else 1e-6
    for ndim in (1, 2, 3):
        nm = tuple(int(v) for v in (75 // ndim + 5 * np.arange(ndim)))
        M = 500
        x = rng.uniform(-np.pi, np.pi, size=(ndim, M)).astype(np.float64 if x64 else np.float32)
        c = (rng.normal(size=M) + 1j * rng.normal(size=M))
        for iflag in (1, -1):
            try:
                f = J.nufft1(nm, T(c, cd), *[T(xx, rd) for xx in x], eps=eps, iflag=iflag).cpu().numpy()
                fd = oracle.dirft1(nm[::-1], c.astype(np.complex64 if not x64 else np.complex128), *x[::-1].astype(np.float64), iflag=iflag)
                e1 = oracle.relerr(f, fd)
                ff = rng.normal(size=nm) + 1j * rng.normal(size=nm)
                c2 = J.nufft2(T(ff, cd), *[T(xx, rd) for xx in x], eps=eps, iflag=iflag).cpu().numpy()
                e2 = oracle.relerr(c2, oracle.dirft2(ff.astype(np.complex64 if not x64 else np.complex128), *x[::-1].astype(np.float64), iflag=iflag))
                s = rng.uniform(-30, 30, size=(ndim, 400)).astype(x.dtype)
                f3 = J.nufft3(T(c, cd), *[T(xx, rd) for xx in x], *[T(ss, rd) for ss in s], eps=eps, iflag=iflag).cpu().numpy()
                e3 = oracle.relerr(f3, oracle.dirft3(c.astype(np.complex64 if not x64 else np.complex128), list(x.astype(np.float64)), list(s.astype(np.float64)), iflag=iflag))
                print(f"x64={x64} ndim={ndim} iflag={iflag:+d}: t1 {e1:.2e} t2 {e2:.2e} t3 {e3:.2e}")
            except Exception as ex:
                print(f"x64={x64} ndim={ndim} iflag={iflag:+d}: FAILED {type(ex).__name__}: {ex}")

print("== mid-size vs oracle(prec=1) and reference cuFINUFFT")
have_ref = ref.available()
for (ndim, nm, M, eps) in ((3, (48, 40, 36), 200000, 1e-6), (2, (200, 180), 200000, 1e-5), (3, (64, 64, 64), 500000, 1e-4)):
    x = rng.uniform(-np.pi, np.pi, size=(ndim, M)).astype(np.float32)
    c = (rng.uniform(-1, 1, size=M) + 1j * rng.uniform(-1, 1, size=M)).astype(np.complex64)
    xt = [T(xx, torch.float32) for xx in x]
    p = Plan(1, nm, eps=eps).setpts(*xt)
    f = p.execute(T(c, torch.complex64)[None])[0].cpu().numpy()
    fo = oracle.nufft1(nm, c, *x.astype(np.float64), eps=eps, prec=1)
    msg = f"t1 ndim={ndim} nm={nm} M={M} eps={eps}: ours-oracle {oracle.relerr(f, fo):.2e}"
    if have_ref:
        r = ref.RefPlan(1, nm, eps=eps).setpts(*xt)
        fr = r.execute(T(c, torch.complex64)[None])[0].cpu().numpy()
        msg += f" ours-ref {oracle.relerr(f, fr):.2e} ref-oracle {oracle.relerr(fr, fo):.2e}"
        r.destroy()
    print(msg)
    ff = (rng.uniform(-1, 1, size=nm[::-1]) + 1j * rng.uniform(-1, 1, size=nm[::-1])).astype(np.complex64)
    p2 = Plan(2, nm, eps=eps).setpts(*xt)
    c2 = p2.execute(T(ff, torch.complex64)[None])[0].cpu().numpy()
    co = oracle.nufft2(ff, *x.astype(np.float64), eps=eps, prec=1)
    msg = f"t2 ndim={ndim}: ours-oracle {oracle.relerr(c2, co):.2e}"
    if have_ref:
        r = ref.RefPlan(2, nm, eps=eps).setpts(*xt)
        cr = r.execute(T(ff, torch.complex64)[None])[0].cpu().numpy()
        msg += f" ours-ref {oracle.relerr(c2, cr):.2e} ref-oracle {oracle.relerr(cr, co):.2e}"
        r.destroy()
    print(msg)
    p.destroy(); p2.destroy()

print("== timing")
def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
for (name, nm, M, eps, typ) in (("C3 3D t1 1e7", (256, 256, 256), 10**7, 1e-6, 1), ("C3 3D t1 1e8", (256, 256, 256), 10**8, 1e-6, 1),
                                 ("C3 3D t2 1e8", (256, 256, 256), 10**8, 1e-6, 2), ("C2 2D t2 1e7", (2048, 2048), 10**7, 1e-5, 2),
                                 ("2D t1 1e7 1024^2", (1024, 1024), 10**7, 1e-6, 1)):
    g = torch.Generator(device=dev); g.manual_seed(1)
    ndim = len(nm)
    xt = [(torch.rand(M, device=dev, generator=g) * 2 - 1) * np.pi for _ in range(ndim)]
    if typ == 1:
        d = torch.complex(torch.rand(M, device=dev, generator=g) * 2 - 1, torch.rand(M, device=dev, generator=g) * 2 - 1)[None]
    else:
        d = torch.complex(torch.rand(nm[::-1], device=dev, generator=g) * 2 - 1, torch.rand(nm[::-1], device=dev, generator=g) * 2 - 1)[None]
    for impl in ("ours", "ref"):
        if impl == "ref" and not have_ref: continue
        try:
            p = Plan(typ, nm, eps=eps) if impl == "ours" else ref.RefPlan(typ, nm, eps=eps)
            ts = timeit(lambda: p.setpts(*xt))
            te = timeit(lambda: p.execute(d))
            print(f"{name} [{impl}]: setpts {ts:.2f} ms  exec {te:.2f} ms  -> {M/(ts+te)*1e3:.3e} pts/s (exec only {M/te*1e3:.3e})")
            if impl == "ours":
                pd = Plan(typ, nm, eps=eps, debug=1); pd.setpts(*xt); pd.execute(d); pd.timings(); pd.setpts(*xt); pd.execute(d); print("   stages:", {k: round(v, 3) for k, v in pd.timings().items()}); pd.destroy()
            p.destroy()
        except Exception as ex:
            print(f"{name} [{impl}] FAILED: {type(ex).__name__}: {ex}")
    del xt, d
    torch.cuda.empty_cache()
