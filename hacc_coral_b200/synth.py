"""Synthetic particle snapshots in the grid units the force tree sees.

All generators return a dict with the reference's ten particle arrays (src/cpu/Particles.h:195-205):
x y z vx vy vz mass phi (float32), id (int64), mask (uint16).  mass = 1 and v = 0 as in the parity runs
(Particles::map2 resets mass to 1 before every kick, src/cpu/Particles.cxx:1256-1257).
"""
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def _pack(x, y, z):
    n = x.size
    z32 = lambda: np.zeros(n, dtype=np.float32)
    return {"x": np.ascontiguousarray(x, dtype=np.float32), "y": np.ascontiguousarray(y, dtype=np.float32),
            "z": np.ascontiguousarray(z, dtype=np.float32), "vx": z32(), "vy": z32(), "vz": z32(),
            "mass": np.ones(n, dtype=np.float32), "phi": z32(), "id": np.arange(n, dtype=np.int64),
            "mask": np.zeros(n, dtype=np.uint16)}


def jitter_lattice(n, seed=1234, amp=0.3):
    """n^3 particles at (i+0.5) + amp*(u-0.5): the survey's near-uniform probe."""
    rng = np.random.default_rng(seed)
    g = np.arange(n, dtype=np.float64) + 0.5
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    u = rng.random((3, n, n, n))
    return _pack((X + amp * (u[0] - 0.5)).ravel(), (Y + amp * (u[1] - 0.5)).ravel(), (Z + amp * (u[2] - 0.5)).ravel())


def clustered(n_particles, box, seed=4321, n_clumps=20, frac=0.5, sigma0=0.3):
    """Half uniform, half in Gaussian clumps of width sigma0*(1+3u) cells (SURVEY.md 8(c) clustered probe)."""
    rng = np.random.default_rng(seed)
    nu = int(n_particles * (1.0 - frac))
    nc = n_particles - nu
    pu = rng.random((nu, 3)) * box
    centers = rng.random((n_clumps, 3)) * (box * 0.8) + box * 0.1
    sig = sigma0 * (1.0 + 3.0 * rng.random(n_clumps))
    which = rng.integers(0, n_clumps, nc)
    pc = centers[which] + rng.standard_normal((nc, 3)) * sig[which, None]
    p = np.concatenate([pu, pc], axis=0)
    p = np.clip(p, 1e-3, box - 1e-3)
    return _pack(p[:, 0], p[:, 1], p[:, 2])


def random_sphere(n_sphere, center, radius, seed=1):
    """Uniform random sphere, the input of the reference's ForceTreeTest (src/halo_finder/ForceTreeTest.cxx:107-140)."""
    rng = np.random.default_rng(seed)
    pts = np.empty((0, 3))
    while pts.shape[0] < n_sphere:
        c = rng.random((2 * n_sphere + 16, 3)) * 2.0 - 1.0
        c = c[(c * c).sum(axis=1) <= 1.0]
        pts = np.concatenate([pts, c], axis=0)
    pts = pts[:n_sphere] * radius + np.asarray(center, dtype=np.float64)
    return _pack(pts[:, 0], pts[:, 1], pts[:, 2])


# ---- Zel'dovich-displaced LambdaCDM particles (the benchmark workload) ----------------------------------
# cosmology of the shipped indat (reference indat:13,23-27): h=0.7, Omega_dm=0.23387755, Omega_b h^2=0.0226,
# n_s=0.97, sigma_8=0.8, box = 0.7875 Mpc/h per particle spacing; z_in = 50 (indat:16).
COSMO = dict(h=0.7, omega_dm=0.23387755, omega_b=0.0226 / 0.7 ** 2, ns=0.97, sigma8=0.8, spacing=0.7875)


def load_transfer():
    """(k [h/Mpc], T(k)) combined CDM+baryon transfer function normalised to T(k->0) = 1.
    Derived from the reference's cmbM000.tf by tests/golden/make_transfer_table.py (the generating script
    is committed; the reference file itself is not copied)."""
    d = np.load(os.path.join(_HERE, "..", "tests", "golden", "transfer_cmbM000.npz"))
    return d["k"], d["T"]


def _growth(a, om, ol):
    """Linear growth factor D(a) (normalised D(1) = 1) for flat LCDM by quadrature of the Heath integral."""
    def E(x):
        return np.sqrt(om / x ** 3 + ol)

    def D(aa):
        xs = np.linspace(1e-6, aa, 4097)
        f = 1.0 / (xs * E(xs)) ** 3
        return E(aa) * np.trapezoid(f, xs)
    return D(a) / D(1.0)


def zeldovich(ns, z=50.0, seed=5009888, ghost=11, growth_boost=1.0, dtype=np.float32):
    """ns^3 particles on a lattice displaced by the Zel'dovich approximation at redshift z, plus a periodic
    ghost shell of `ghost` cells on every face (the overload zone, reference src/simulation/Domain.cxx:53-79),
    shifted so that all coordinates lie in [0, ns + 2*ghost).  Displacements: x = q - D(z) grad(phi)
    (reference src/initializer/Initializer.cpp:1141-1142) with a Gaussian field of spectrum
    P(k) = A k^ns T(k)^2 normalised to sigma_8.  growth_boost > 1 pushes the field into the clustered
    (shell-crossed) regime used for the 'evolved' benchmark configuration."""
    import numpy.fft as fft
    c = COSMO
    L = ns * c["spacing"]                       # Mpc/h
    kf = 2.0 * np.pi / L
    k1 = fft.fftfreq(ns, d=1.0 / ns) * kf
    kz = fft.rfftfreq(ns, d=1.0 / ns) * kf
    KX, KY, KZ = np.meshgrid(k1, k1, kz, indexing="ij")
    K2 = KX ** 2 + KY ** 2 + KZ ** 2
    K = np.sqrt(K2)
    kt, Tt = load_transfer()
    T = np.interp(np.log(np.maximum(K, kt[0])), np.log(kt), Tt)
    P = np.where(K > 0, np.maximum(K, 1e-30) ** c["ns"] * T ** 2, 0.0)
    # sigma_8 normalisation by direct quadrature of the continuous spectrum
    kk = np.logspace(np.log10(kt[0]), np.log10(kt[-1]), 4000)
    Tk = np.interp(np.log(kk), np.log(kt), Tt)
    x = kk * 8.0
    W = 3.0 * (np.sin(x) - x * np.cos(x)) / x ** 3
    s2 = np.trapezoid(kk ** (2 + c["ns"]) * Tk ** 2 * W ** 2 / (2.0 * np.pi ** 2), kk)
    A = c["sigma8"] ** 2 / s2
    rng = np.random.default_rng(seed)
    white = rng.standard_normal((ns, ns, ns))
    # numpy's unnormalised forward FFT of unit white noise has <|W_k|^2> = N^3; the field delta(x) =
    # irfftn(F) has the spectrum P when <|F_k|^2> = N^6 P / V.
    dk = fft.rfftn(white) * np.sqrt(A * P * float(ns) ** 3 / L ** 3)
    del white
    om = c["omega_dm"] + c["omega_b"]
    D = _growth(1.0 / (1.0 + z), om, 1.0 - om) * growth_boost
    with np.errstate(divide="ignore", invalid="ignore"):
        invk2 = np.where(K2 > 0, 1.0 / np.maximum(K2, 1e-300), 0.0)
    disp = []
    for Kc in (KX, KY, KZ):
        # psi_k = i k delta_k / k^2  (div psi = -delta), i.e. x = q - D grad(phi) with laplacian(phi) = delta
        psi = fft.irfftn(1j * Kc * dk * invk2, s=(ns, ns, ns), axes=(0, 1, 2))
        disp.append(D * psi / c["spacing"])                      # Mpc/h -> grid units
    g = np.arange(ns, dtype=np.float64) + 0.5
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    pos = [np.mod(X + disp[0], ns), np.mod(Y + disp[1], ns), np.mod(Z + disp[2], ns)]
    pos = np.stack([p.ravel() for p in pos], axis=1)
    if ghost > 0:
        # periodic images that fall in the overload shell [-ghost, ns+ghost)
        out = []
        for sx in (-1, 0, 1):
            for sy in (-1, 0, 1):
                for sz in (-1, 0, 1):
                    q = pos + np.array([sx, sy, sz], dtype=np.float64) * ns
                    m = np.all((q >= -ghost) & (q < ns + ghost), axis=1)
                    out.append(q[m])
        pos = np.concatenate(out, axis=0) + ghost
    pos = pos.astype(dtype)
    # keep strictly inside the tree box after the float32 cast
    hi = np.float32(ns + 2 * ghost)
    pos = np.minimum(pos, np.nextafter(hi, np.float32(0)))
    return _pack(pos[:, 0], pos[:, 1], pos[:, 2])


def zeldovich_torch(ns, z=50.0, seed=5009888, ghost=11, growth_boost=1.0, device="cuda", return_tensor=False):
    """Same recipe as zeldovich() evaluated with torch FFTs on `device` (used by bench.py so that a
    256^3 snapshot takes seconds).  Returns the usual dict of numpy arrays (host).  `ns` may be a triple (nx, ny, nz):
    a periodic box of nx x ny x nz particles at the same spacing (the global box of a 2x1x1 or 2x2x1 decomposition of
    equal sub-volumes); only with return_tensor."""
    import torch
    c = COSMO
    shape = (int(ns),) * 3 if np.isscalar(ns) else tuple(int(t) for t in ns)
    assert return_tensor or shape[0] == shape[1] == shape[2]
    dev = torch.device(device)
    ks = []
    for ax in range(3):
        L = shape[ax] * c["spacing"]
        f = torch.fft.rfftfreq if ax == 2 else torch.fft.fftfreq
        ks.append(f(shape[ax], d=1.0 / shape[ax], device=dev, dtype=torch.float64) * (2.0 * np.pi / L))
    KX, KY, KZ = torch.meshgrid(ks[0], ks[1], ks[2], indexing="ij")
    K2 = KX ** 2 + KY ** 2 + KZ ** 2
    kt, Tt = load_transfer()
    K = torch.sqrt(K2).clamp_min(float(kt[0]))
    # log-log interpolation of T(k) with torch.searchsorted
    lk = torch.log(K)
    lkt = torch.as_tensor(np.log(kt), device=dev, dtype=torch.float64)
    Ttt = torch.as_tensor(Tt, device=dev, dtype=torch.float64)
    idx = torch.searchsorted(lkt, lk.reshape(-1)).clamp(1, lkt.numel() - 1).reshape(lk.shape)
    w = (lk - lkt[idx - 1]) / (lkt[idx] - lkt[idx - 1])
    T = Ttt[idx - 1] * (1 - w) + Ttt[idx] * w
    P = torch.where(K2 > 0, K ** c["ns"] * T ** 2, torch.zeros_like(K2))
    kk = np.logspace(np.log10(kt[0]), np.log10(kt[-1]), 4000)
    Tk = np.interp(np.log(kk), np.log(kt), Tt)
    x = kk * 8.0
    W = 3.0 * (np.sin(x) - x * np.cos(x)) / x ** 3
    s2 = np.trapezoid(kk ** (2 + c["ns"]) * Tk ** 2 * W ** 2 / (2.0 * np.pi ** 2), kk)
    A = c["sigma8"] ** 2 / s2
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed))
    white = torch.randn(shape, generator=gen, device=dev, dtype=torch.float64)
    # particles per volume: N / V = spacing^-3 whatever the shape of the box
    dk = torch.fft.rfftn(white) * torch.sqrt(A * P / c["spacing"] ** 3)
    del white
    om = c["omega_dm"] + c["omega_b"]
    D = _growth(1.0 / (1.0 + z), om, 1.0 - om) * growth_boost
    invk2 = torch.where(K2 > 0, 1.0 / K2.clamp_min(1e-300), torch.zeros_like(K2))
    pos = []
    for ax, Kc in enumerate((KX, KY, KZ)):
        psi = torch.fft.irfftn(1j * Kc * dk * invk2, s=shape)
        sh = [1, 1, 1]
        sh[ax] = shape[ax]
        g = torch.arange(shape[ax], device=dev, dtype=torch.float64) + 0.5
        pos.append(torch.remainder(g.reshape(sh) + D * psi / c["spacing"], shape[ax]).reshape(-1))
    pos = torch.stack(pos, dim=1)
    if return_tensor:          # (nx*ny*nz, 3) float64 positions in [0, n_axis) on `device`, no ghost shell
        return pos
    ns = shape[0]
    if ghost > 0:
        out = []
        for sx in (-1, 0, 1):
            for sy in (-1, 0, 1):
                for sz in (-1, 0, 1):
                    q = pos + torch.tensor([sx, sy, sz], device=dev, dtype=torch.float64) * ns
                    m = ((q >= -ghost) & (q < ns + ghost)).all(dim=1)
                    out.append(q[m])
        pos = torch.cat(out, dim=0) + ghost
    hi = np.nextafter(np.float32(ns + 2 * ghost), np.float32(0))
    pos = pos.to(torch.float32).clamp_max(float(hi)).cpu().numpy()
    return _pack(pos[:, 0], pos[:, 1], pos[:, 2])


def add_clumps(p, side, frac=0.15, n_clumps=64, seed=99, r_lo=0.4, r_hi=2.5):
    """Move a fraction of the particles of snapshot p (coordinates in [0, side)) into isothermal (rho ~ r^-2) clumps of
    radius r_lo..r_hi cells at seeded random centres: halo-like knots of a few 10^4 particles with overdensities of
    10^3-10^5, which produce the deep interaction lists of an evolved snapshot (state C).  In place."""
    rng = np.random.default_rng(seed)
    n = p["x"].size
    pick = rng.choice(n, int(frac * n), replace=False)
    centres = rng.random((n_clumps, 3)) * (side - 8.0) + 4.0
    radius = r_lo + (r_hi - r_lo) * rng.random(n_clumps)
    which = rng.integers(0, n_clumps, pick.size)
    r = radius[which] * rng.random(pick.size)                    # uniform in r  =>  rho ~ r^-2
    u = rng.standard_normal((pick.size, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    q = centres[which] + r[:, None] * u
    top = np.nextafter(np.float32(side), np.float32(0))
    for k, a in enumerate(("x", "y", "z")):
        p[a][pick] = np.clip(q[:, k], 0.0, top).astype(np.float32)
    return p


def cutout(p, lo, hi):
    """Particles of snapshot p inside the cube [lo, hi)^3, shifted to start at 0 (a bounded sample of the
    same workload for the CPU baseline)."""
    m = np.ones(p["x"].size, dtype=bool)
    for k in ("x", "y", "z"):
        m &= (p[k] >= lo) & (p[k] < hi)
    q = {k: np.ascontiguousarray(v[m]) for k, v in p.items()}
    for k in ("x", "y", "z"):
        q[k] = (q[k] - np.float32(lo)).astype(np.float32)
    q["id"] = np.arange(q["x"].size, dtype=np.int64)
    return q
