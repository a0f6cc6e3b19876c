"""Host models of device-side index logic, checked against the oracle on the CPU.

These are line-by-line Python restatements of the enumeration a CUDA kernel performs (which
thread owns which (ray, voxel) pair, which samples it integrates over, which quadrature weight each
sample gets).  They are test infrastructure: they let the no-GPU suite catch a wrong ownership rule
or weight formula before the kernel ever runs; the kernel itself is compared with the oracle in
tests/test_gpu_parity.py."""
import numpy as np

from oracle import ionotomo_oracle as O


def simpson_weight_model(i, N, s):
    """``simpson_weight`` of csrc/iono_device.cuh with its neighbour clamping, exact division."""
    c = lambda k: s[min(max(k, 0), N - 1)]
    sm2, sm1, s0, sp1, sp2 = c(i - 2), c(i - 1), c(i), c(i + 1), c(i + 2)
    n_odd = N % 2 == 1
    hm, hp = s0 - sm1, sp1 - s0
    if i < 1:
        hm = hp
    if i > N - 2:
        hp = hm
    h0 = sm1 - sm2 if i >= 2 else hm
    h3 = sp2 - sp1 if i <= N - 3 else hp
    par = i & 1
    trap = 0.0
    if n_odd:
        cm = 1.0 if par else 0.0
        cr = 0.0 if (par or i < 2) else 1.0
        cl = 0.0 if (par or i > N - 3) else 1.0
    else:
        cm = 0.5 if ((i <= N - 3) if par else (i >= 2)) else 0.0
        cr = 0.5 if i >= 2 + par else 0.0
        cl = 0.5 if i <= N - 4 + par else 0.0
        if i < 2 or i > N - 3:
            trap = 0.25 * (hp * ((i == 0) + (i == N - 2)) + hm * ((i == 1) + (i == N - 1)))
    A = hm + hp
    num = cm * A * A * A + cr * (h0 + hm) * (2.0 * hm - h0) * hp + cl * (hp + h3) * (2.0 * hp - h3) * hm
    return num / (6.0 * hm * hp) + trap


def test_simpson_weight_model_small_segments():
    """Every segment length the Gaussian adjoint can meet, 2 upwards, non-uniform abscissae."""
    rng = np.random.RandomState(0)
    for N in range(2, 12):
        s = np.cumsum(rng.uniform(0.5, 1.5, N))
        w = np.array([simpson_weight_model(i, N, s) for i in range(N)])
        np.testing.assert_allclose(w, O.simps_weights(s), rtol=1e-12, atol=1e-14)
        y = rng.normal(size=N)
        np.testing.assert_allclose(w @ y, O.simps_avg(y, s), rtol=1e-11, atol=1e-13)


def gaussian_adjoint_kernel_model(ray, ne_ray, xvec, yvec, zvec, sigma_m, L_m, Nk):
    """gaussian_adjoint_kernel (csrc/iono_gaussian.cuh), one ray, lanes run one after another."""
    x, y, z, s = ray
    Ns = len(x)
    nx, ny, nz = len(xvec), len(yvec), len(zvec)
    cx = [O.bisection(xvec, v) for v in x]
    cy = [O.bisection(yvec, v) for v in y]
    cz = [O.bisection(zvec, v) for v in z]
    mono = all(cz[q] >= cz[q - 1] for q in range(1, Ns))

    def member(q, xi, yi, zi):
        return abs(cx[q] - xi) <= Nk and abs(cy[q] - yi) <= Nk and abs(cz[q] - zi) <= Nk

    out = {}
    for zi in range(0, nz - 1):
        s_lo, s_hi = 0, Ns
        if mono:
            lo, hi = 0, Ns
            while lo < hi:
                m = (lo + hi) >> 1
                if cz[m] < zi - Nk:
                    lo = m + 1
                else:
                    hi = m
            s_lo = lo
            hi = Ns
            while lo < hi:
                m = (lo + hi) >> 1
                if cz[m] <= zi + Nk:
                    lo = m + 1
                else:
                    hi = m
            s_hi = lo
        for q0 in range(s_lo, s_hi):
            if abs(cz[q0] - zi) > Nk:
                continue
            for xi in range(max(0, cx[q0] - Nk), min(nx - 2, cx[q0] + Nk) + 1):
                for yi in range(max(0, cy[q0] - Nk), min(ny - 2, cy[q0] + Nk) + 1):
                    if any(member(q, xi, yi, zi) for q in range(s_lo, q0)):
                        continue
                    b = q0
                    for q in range(q0 + 1, s_hi):
                        if member(q, xi, yi, zi):
                            b = q
                    n = b - q0 + 1
                    if n < 2:
                        continue
                    total = 0.0
                    for i in range(n):
                        q = q0 + i
                        r2 = (xvec[xi] - x[q]) ** 2 + (yvec[yi] - y[q]) ** 2 + (zvec[zi] - z[q]) ** 2
                        f = np.exp(r2 / (-2.0 * L_m * L_m)) * sigma_m ** 2 * ne_ray[q]
                        total += simpson_weight_model(i, n, s[q0:b + 1]) * f
                    assert (xi, yi, zi) not in out          # each (ray, voxel) pair has exactly one owner
                    out[(xi, yi, zi)] = total
    return out


def _compare(ray, ne_ray, xv, yv, zv, sigma_m, L_m, Nk):
    ref = O.gaussian_adjoint_ray(ray, ne_ray, xv, yv, zv, sigma_m, L_m, Nk)
    got = gaussian_adjoint_kernel_model(ray, ne_ray, xv, yv, zv, sigma_m, L_m, Nk)
    scale = max(abs(v) for v in ref.values())
    for v, c in ref.items():                                   # the oracle also lists one-sample segments (= 0)
        assert abs(got.get(v, 0.0) - c) <= 1e-11 * scale, (v, got.get(v), c)
    assert set(got) <= set(ref)
    return len(got)


def test_gaussian_adjoint_kernel_model_on_reference_rays(golden):
    g = golden("adjoint_gauss")
    xv, yv, zv = g["xvec"], g["yvec"], g["zvec"]
    rng = np.random.RandomState(1)
    n = 0
    for (i, j, k), Nk in (((0, 0, 0), 2), ((2, 1, 1), 2), ((1, 0, 1), 1), ((2, 0, 0), 5), ((0, 1, 0), 0)):
        ray = g["rays"][i, j, k]
        n += _compare(ray, rng.uniform(0.5, 2., ray.shape[1]), xv, yv, zv, 0.7, 20. * max(Nk, 1), Nk)
    assert n > 500


def test_gaussian_adjoint_kernel_model_non_monotone_ray(golden):
    """A ray that turns back in z: first/last-sample hull semantics, all samples scanned."""
    g = golden("adjoint_gauss")
    xv, yv, zv = g["xvec"], g["yvec"], g["zvec"]
    t = np.linspace(0., 1., 17)
    ray = np.stack([-30. + 70. * t, 20. * np.sin(5. * t), 500. + 450. * np.sin(7. * t),
                    np.cumsum(np.full(17, 60.)) + 3. * np.cos(11. * t)])
    assert np.any(np.diff(ray[2]) < 0)
    assert _compare(ray, np.linspace(1., 2., 17), xv, yv, zv, 1.1, 35., 2) > 100


# ---------------------------------------------------------------- run-compressed ray indices
def run_records_model(idx, seg=256):
    """run_record_units_kernel / run_record_fill_kernel (csrc/iono_backproject.cuh): per segment of `seg`
    entries one run number per entry, stored lane-major (byte lane*8 + u <-> entry 32u + lane), and per run
    the base = ray index of its head - position of the head (mod 2^32)."""
    recs = []
    for k0 in range(0, len(idx), seg):
        ids, bases = bytearray(seg), []
        for u in range(seg // 32):
            for lane in range(32):
                j = 32 * u + lane
                if j == 0 or idx[k0 + j] != (idx[k0 + j - 1] + 1) & 0xffffffff:
                    bases.append((int(idx[k0 + j]) - j) & 0xffffffff)
                ids[lane * 8 + u] = len(bases) - 1
        recs.append((bytes(ids), bases))
    return recs


def run_index_model(ids, bases, u, lane):
    """Ray index of entry 32u+lane as backproject_wruns_kernel reconstructs it: the lane's 8 run numbers are one
    64-bit shared-memory load (two 32-bit words), byte u of them selects the base."""
    lo = int.from_bytes(ids[lane * 8:lane * 8 + 4], "little")
    hi = int.from_bytes(ids[lane * 8 + 4:lane * 8 + 8], "little")
    run = ((lo if u < 4 else hi) >> (8 * (u & 3))) & 0xff
    return (bases[run] + 32 * u + lane) & 0xffffffff


def test_run_compressed_indices_model():
    rng = np.random.RandomState(3)
    # rows of a voxel-sorted operator: runs of consecutive rays (lengths 1..40), restarting at every row,
    # plus the zero padding of the last segment and a run that crosses a segment boundary
    idx = []
    while len(idx) < 5 * 256 - 37:
        start = rng.randint(0, 1_000_000)
        for _ in range(rng.randint(1, 12)):
            n = rng.randint(1, 41)
            idx.extend(range(start, start + n))
            start += n + rng.randint(1, 50_000)
    idx = np.array(idx[:5 * 256 - 37] + [0] * 37, dtype=np.uint32)
    recs = run_records_model(idx)
    assert len(recs) == 5
    n_heads = 0
    for g, (ids, bases) in enumerate(recs):
        assert 1 <= len(bases) <= 256                        # run numbers fit a byte
        n_heads += len(bases)
        for u in range(8):
            for lane in range(32):
                assert run_index_model(ids, bases, u, lane) == idx[g * 256 + 32 * u + lane]
    assert n_heads < len(idx) / 4                            # the point of the format
    # record size in 16-byte units as the build computes it
    assert all((256 + 4 * len(b) + 15) // 16 * 16 <= 256 + 256 * 4 for _, b in recs)


# ---------------------------------------------------------------- face bookkeeping of the prepared adjoint
def face_walk_model(cells, contribs, sx, sy, grid_size):
    """One lane-sample of ``prepared_adjoint_kernel`` (csrc/iono_prepared_adjoint.cuh, lambda ``leave``): walk the
    time axis with the 8 corner sums ``a[4x + 2y + z]`` of the current cell in registers; on a change of cell queue
    only the FACE that is finished -- (base, stride; 4 values) reduced at base, base+1, base+stride, base+stride+1 --
    and carry the shared face over.  Returns the accumulator and the number of faces flushed."""
    acc = np.zeros(grid_size)
    faces = 0

    def flush(base, stride, f):
        nonlocal faces
        faces += 1
        for off, val in zip((0, 1, stride, stride + 1), f):
            acc[base + off] += val

    a = [0.0] * 8
    vc = -1
    for v, l in zip(cells, contribs):
        if vc >= 0 and v != vc:
            delta = v - vc
            dx = 1 if abs(delta - sx) <= sy else (-1 if abs(delta + sx) <= sy else 0)
            ry = delta - dx * sx
            dy = 1 if ry == sy else (-1 if ry == -sy else 0)
            jump = ry != dy * sy
            # round A
            if jump:
                flush(vc, sy, a[0:4]); a[0:4] = [0.0] * 4
            elif dx > 0:
                flush(vc, sy, a[0:4]); a[0:4] = a[4:8]; a[4:8] = [0.0] * 4
            elif dx < 0:
                flush(vc + sx, sy, a[4:8]); a[4:8] = a[0:4]; a[0:4] = [0.0] * 4
            # round B
            vm = vc + dx * sx
            if jump:
                flush(vc + sx, sy, a[4:8]); a[4:8] = [0.0] * 4
            elif dy > 0:
                flush(vm, sx, [a[0], a[1], a[4], a[5]])
                a[0], a[1], a[4], a[5] = a[2], a[3], a[6], a[7]
                a[2] = a[3] = a[6] = a[7] = 0.0
            elif dy < 0:
                flush(vm + sy, sx, [a[2], a[3], a[6], a[7]])
                a[2], a[3], a[6], a[7] = a[0], a[1], a[4], a[5]
                a[0] = a[1] = a[4] = a[5] = 0.0
        vc = v
        a = [ai + li for ai, li in zip(a, l)]
    if vc >= 0:                                   # end of the task: the cell is left for good
        flush(vc, sy, a[0:4])
        flush(vc + sx, sy, a[4:8])
    return acc, faces


def test_face_walk_equals_direct_accumulation():
    """Random walks with face moves, diagonal moves, z moves and jumps: the face bookkeeping deposits exactly what
    adding every step's 8 corner contributions directly would (up to the order of the additions), with about one
    face per move instead of two per cell."""
    rng = np.random.RandomState(11)
    nx, ny, nz = 9, 8, 7
    sy, sx = nz, ny * nz
    for trial in range(40):
        ix, iy, iz = rng.randint(2, nx - 3), rng.randint(2, ny - 3), rng.randint(1, nz - 2)
        cells, contribs = [], []
        for t in range(60):
            r = rng.rand()
            if r < 0.25:
                ix += rng.choice([-1, 1])
            elif r < 0.5:
                iy += rng.choice([-1, 1])
            elif r < 0.6:
                ix += rng.choice([-1, 1]); iy += rng.choice([-1, 1])
            elif r < 0.65:
                iz += rng.choice([-1, 1])
            elif r < 0.7:
                ix, iy = rng.randint(0, nx - 1), rng.randint(0, ny - 1)
            ix, iy, iz = int(np.clip(ix, 0, nx - 2)), int(np.clip(iy, 0, ny - 2)), int(np.clip(iz, 0, nz - 2))
            cells.append((ix * ny + iy) * nz + iz)
            contribs.append(list(rng.normal(size=8)))
        acc, faces = face_walk_model(cells, contribs, sx, sy, nx * ny * nz)
        ref = np.zeros(nx * ny * nz)
        for v, l in zip(cells, contribs):
            for e in range(8):
                ref[v + (e >> 2) * sx + ((e >> 1) & 1) * sy + (e & 1)] += l[e]
        np.testing.assert_allclose(acc, ref, rtol=0, atol=1e-12)
        n_cells = 1 + sum(1 for p, q in zip(cells[:-1], cells[1:]) if p != q)
        assert faces <= 2 * n_cells


# ---------------------------------------------------------------- factored Simpson weights of the prepared operator
def weights_factor_model(S, tol=2e-13):
    """``weight_pattern_kernel`` + ``weight_factor_kernel`` of csrc/iono_prepared.cuh: pattern from the first ray
    (``P = w0 * Ns / sum(w0)``), per-ray factor ``c = sum(w) / Ns``, accepted when ``|w - c P| <= tol |c P|``
    for every sample of every ray.  ``S``: (R, Ns) abscissae.  Returns ``(factored, P, c)``."""
    R, Ns = S.shape
    W = np.array([[simpson_weight_model(i, Ns, s) for i in range(Ns)] for s in S])
    P = W[0] * Ns / W[0].sum()
    c = W.sum(axis=1) / Ns
    want = c[:, None] * P[None, :]
    return bool(np.all(np.abs(W - want) <= tol * np.abs(want))), P, c


def test_weight_factoring_accepts_linspace_and_rejects_perturbed_abscissae():
    """Rays made by the casting kernels (s a linspace of ray-dependent length, rounded) factor into a common
    pattern x a per-ray length; the factored operator then reproduces the oracle's Simpson integral to rounding.
    Non-uniform abscissae that differ from ray to ray do not factor (the operator keeps per-sample weights)."""
    rng = np.random.RandomState(3)
    for Ns in (2, 3, 4, 5, 30, 31, 64, 128):
        smax = rng.uniform(800., 1400., size=6)
        S = np.array([(np.linspace(-10., 1000., Ns) - (-10.)) / (1010. / sm) for sm in smax])   # (z - z0)/pz: rounded
        ok, P, c = weights_factor_model(S)
        assert ok, Ns
        y = rng.normal(size=S.shape)
        got = c * (P[None, :] * y).sum(axis=1)
        ref = np.array([O.simps_avg(y[r], S[r]) for r in range(S.shape[0])]) if hasattr(O, "simps_avg") else \
            (np.array([[simpson_weight_model(i, Ns, S[r]) for i in range(Ns)] for r in range(S.shape[0])]) * y).sum(axis=1)
        np.testing.assert_allclose(got, ref, rtol=0, atol=2e-12 * np.abs(y).max() * smax.max())
        if Ns >= 3:
            Sp = S + 0.3 * np.sin(S / 50.)
            assert not weights_factor_model(Sp)[0], Ns
