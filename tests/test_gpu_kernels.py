"""Per-routine parity of the CUDA kernels against the CPU oracle, through the C ABI
(include/lbfgsb_b200.h section 5).  Needs a B200: run with `-m gpu`.

Bit-exact where the arithmetic is order-free or replays a fixed order (max, sort, the 2m x 2m
dense algebra, dcsrch/dcstep, the fixed-shape long sums against the oracle's device-order mode).
"""
import ctypes as C

import numpy as np
import pytest

import harness as H
from oracle import oracle_py as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    import torch
    assert torch.cuda.is_available()
    import lbfgsb_b200
    return lbfgsb_b200.lib()


def _dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _vp(t):
    return C.c_void_p(t.data_ptr())


@pytest.mark.parametrize("n", [1, 2, 3, 31, 255, 2048, 2049, 100003, 1 << 20, 3000017])
def test_fixed_shape_sum_f64_bitwise(L, n):
    rng = np.random.default_rng(n)
    a = rng.standard_normal(n) * np.exp(rng.uniform(-8, 8, n))
    b = rng.standard_normal(n)
    out = np.zeros(1)
    assert L.lbfgsb_test_sum_f64(C.c_int64(n), _vp(_dev(a)), _vp(_dev(b)), out.ctypes.data_as(C.c_void_p)) == 0
    lo = O.lib()
    lo.oracle_device_order_dot_f64.restype = C.c_double
    ref = lo.oracle_device_order_dot_f64(C.c_int64(n), a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p))
    assert out[0] == ref, (out[0], ref, float(a @ b))


@pytest.mark.parametrize("n", [5, 4097, 777777])
def test_fixed_shape_sum_f32_bitwise(L, n):
    rng = np.random.default_rng(n)
    a = rng.standard_normal(n).astype(np.float32)
    b = rng.standard_normal(n).astype(np.float32)
    out = np.zeros(1, np.float32)
    assert L.lbfgsb_test_sum_f32(C.c_int64(n), _vp(_dev(a)), _vp(_dev(b)), out.ctypes.data_as(C.c_void_p)) == 0
    lo = O.lib()
    lo.oracle_device_order_dot_f32.restype = C.c_float
    ref = lo.oracle_device_order_dot_f32(C.c_int64(n), a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p))
    assert out[0] == np.float32(ref)


@pytest.mark.parametrize("n", [1, 7, 1000, 123457, 2000003])
def test_projgr_exact(L, n):
    """projgr src/lbfgsb.f90:2594-2622 -- a max, so exact whatever the order."""
    rng = np.random.default_rng(n + 1)
    x = rng.uniform(-2, 2, n)
    l = x - rng.uniform(0, 1, n) * (rng.random(n) < 0.7)
    u = x + rng.uniform(0, 1, n) * (rng.random(n) < 0.7)
    nbd = rng.integers(0, 4, n).astype(np.int32)
    g = rng.standard_normal(n) * 3
    out = np.zeros(1)
    assert L.lbfgsb_test_projgr_f64(C.c_int64(n), _vp(_dev(l)), _vp(_dev(u)), _vp(_dev(nbd)), _vp(_dev(x)),
                                    _vp(_dev(g)), out.ctypes.data_as(C.c_void_p)) == 0
    ref = np.zeros(1)
    O.lib().oracle_projgr_f64(C.c_int64(n), l.ctypes.data_as(C.c_void_p), u.ctypes.data_as(C.c_void_p),
                              nbd.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p),
                              g.ctypes.data_as(C.c_void_p), ref.ctypes.data_as(C.c_void_p))
    assert out[0] == ref[0]


@pytest.mark.parametrize("n,ties", [(1, False), (100, False), (5000, True), (1 << 16, False), (1234567, True)])
def test_breakpoint_sort_is_stable(L, n, ties):
    """Replaces hpsolb (src/lbfgsb.f90:2079-2157): ascending t, ties in index order."""
    import torch
    rng = np.random.default_rng(n)
    t = rng.uniform(1e-12, 10.0, n) * np.exp(rng.uniform(-20, 20, n))
    if ties:
        t = np.round(t, 1) + 0.5        # many exact ties
        t[: n // 3] = 0.0               # breakpoints at t = 0 (variables sitting on a bound with g pointing out)
    td = _dev(t)
    order = torch.empty(n, dtype=torch.int32, device="cuda")
    srt = torch.empty(n, dtype=torch.float64, device="cuda")
    assert L.lbfgsb_test_sort_f64(C.c_int64(n), _vp(td), _vp(order), _vp(srt)) == 0
    ref = np.argsort(t, kind="stable")
    assert np.array_equal(order.cpu().numpy(), ref.astype(np.int32))
    assert np.array_equal(srt.cpu().numpy(), t[ref])


@pytest.mark.parametrize("n,distinct", [(1, 1), (2, 1), (7, 2), (100, 1), (1000, 3), (4097, 17), (50000, 5), (50000, 50000)])
def test_heap_replay_pops_in_hpsolb_order(L, n, distinct):
    """hpsolb (src/lbfgsb.f90:2079-2157) replayed on the device pops equal keys in the reference's order
    (bit-exact against the oracle's verbatim heap)."""
    import torch
    rng = np.random.default_rng(n + distinct)
    vals = rng.uniform(0.1, 5.0, distinct)
    t = vals[rng.integers(0, distinct, n)].astype(np.float64)
    order = torch.empty(n, dtype=torch.int32, device="cuda")
    assert L.lbfgsb_test_heap_order_f64(C.c_int64(n), _vp(_dev(t)), _vp(order)) == 0
    # the oracle: build on the first call (iheap = 0), then pop; the least member is left in t(nleft)
    tt = t.copy()
    io = np.arange(n, dtype=np.int32)
    ref = np.empty(n, dtype=np.int32)
    for k in range(n):
        nleft = n - k
        O.lib().oracle_hpsolb_f64(C.c_int64(nleft), tt.ctypes.data_as(C.c_void_p), io.ctypes.data_as(C.c_void_p),
                                  C.c_int64(0 if k == 0 else 1))
        ref[k] = io[nleft - 1]
    got = order.cpu().numpy()
    assert np.array_equal(got, ref)
    assert np.all(np.diff(t[got]) >= 0)


def _spd(rng, m, col):
    a = rng.standard_normal((col, col))
    s = a @ a.T + col * np.eye(col)
    full = np.zeros((m, m))
    full[:col, :col] = s
    return np.asfortranarray(full)


@pytest.mark.parametrize("m,col", [(5, 1), (5, 5), (10, 7), (20, 20)])
def test_dpofa_dtrsl_bitwise(L, m, col):
    """dpofa / dtrsl, src/lbfgsb_linpack_module.f90:30-67, :87-165."""
    rng = np.random.default_rng(m * 100 + col)
    a = _spd(rng, m, col)
    ref = a.copy(order="F")
    info_ref = np.zeros(1, np.int32)
    O.lib().oracle_dpofa_f64(ref.ctypes.data_as(C.c_void_p), m, col, info_ref.ctypes.data_as(C.c_void_p))
    ad = _dev(a.T)     # column-major bytes
    info = np.zeros(1, np.int32)
    dummy = _dev(np.zeros(4 * m))
    assert L.lbfgsb_test_dense_f64(0, m, col, C.c_double(1.0), _vp(ad), _vp(dummy), _vp(dummy),
                                   info.ctypes.data_as(C.c_void_p)) == 0
    got = ad.cpu().numpy().T
    assert info[0] == info_ref[0] == 0
    assert np.array_equal(np.triu(got[:col, :col]), np.triu(ref[:col, :col]))
    for job, op in ((1, 1), (11, 2)):
        b = rng.standard_normal(col)
        bref = b.copy()
        O.lib().oracle_dtrsl_f64(ref.ctypes.data_as(C.c_void_p), m, col, bref.ctypes.data_as(C.c_void_p), job,
                                 info_ref.ctypes.data_as(C.c_void_p))
        bd = _dev(b)
        assert L.lbfgsb_test_dense_f64(op, m, col, C.c_double(1.0), _vp(ad), _vp(bd), _vp(dummy),
                                       info.ctypes.data_as(C.c_void_p)) == 0
        assert info[0] == info_ref[0] == 0
        assert np.array_equal(bd.cpu().numpy(), bref)


@pytest.mark.parametrize("m,col", [(5, 1), (5, 5), (10, 7), (10, 10), (20, 13), (20, 20)])
@pytest.mark.parametrize("warp", [0, 10])
def test_bmv_formt_formk_tail_bitwise(L, m, col, warp):
    """bmv (src/lbfgsb.f90:1057-1123), formt (:1926-1963) and the dense tail of formk (:1853-1906) against the oracle,
    bit for bit; warp = 0: the single-thread routines, 10: the one-warp versions the scalar kernels run (dpofa, dtrsl
    included), which form independent entries in different lanes but every entry by the same operations in the same order."""
    rng = np.random.default_rng(1000 * m + col + warp)
    lo = O.lib()
    vp = lambda a: a.ctypes.data_as(C.c_void_p)   # noqa: E731
    info_ref = np.zeros(1, np.int32)
    info = np.zeros(1, np.int32)
    # S, Y with S'Y positive diagonal: sy = lower triangle of S'Y incl. diagonal, ss = S'S
    S = rng.standard_normal((4 * m, col))
    Y = S * rng.uniform(0.5, 2.0, (4 * m, 1)) + 0.1 * rng.standard_normal((4 * m, col))
    sy = np.zeros((m, m), order="F"); ss = np.zeros((m, m), order="F")
    sy[:col, :col] = S.T @ Y; ss[:col, :col] = S.T @ S
    theta = float((Y[:, -1] @ Y[:, -1]) / (S[:, -1] @ Y[:, -1]))
    # formt
    wt_ref = np.zeros((m, m), order="F")
    lo.oracle_formt_f64(m, vp(wt_ref), vp(sy), vp(ss), col, C.c_double(theta), vp(info_ref))
    syd, ssd, wtd = _dev(sy.T), _dev(ss.T), _dev(np.zeros((m, m)))
    assert L.lbfgsb_test_dense_f64(4 + warp, m, col, C.c_double(theta), _vp(syd), _vp(ssd), _vp(wtd), vp(info)) == 0
    assert info[0] == info_ref[0] == 0
    got = wtd.cpu().numpy().T
    assert np.array_equal(np.triu(got[:col, :col]), np.triu(wt_ref[:col, :col]))
    if warp:   # dpofa / dtrsl by one warp on this factor
        a = _spd(rng, m, col); ref = a.copy(order="F")
        lo.oracle_dpofa_f64(vp(ref), m, col, vp(info_ref))
        ad = _dev(a.T); dummy = _dev(np.zeros(4 * m))
        assert L.lbfgsb_test_dense_f64(10, m, col, C.c_double(1.0), _vp(ad), _vp(dummy), _vp(dummy), vp(info)) == 0
        assert info[0] == info_ref[0] == 0
        assert np.array_equal(np.triu(ad.cpu().numpy().T[:col, :col]), np.triu(ref[:col, :col]))
        for job, op in ((1, 11), (11, 12)):
            b = rng.standard_normal(col); bref = b.copy()
            lo.oracle_dtrsl_f64(vp(ref), m, col, vp(bref), job, vp(info_ref))
            bd = _dev(b)
            assert L.lbfgsb_test_dense_f64(op, m, col, C.c_double(1.0), _vp(ad), _vp(bd), _vp(dummy), vp(info)) == 0
            assert np.array_equal(bd.cpu().numpy(), bref)
    # bmv with that wt
    v = rng.standard_normal(2 * m); v[2 * col:] = 0.0
    vv = np.zeros(2 * col); vv[:] = v[:2 * col]
    p_ref = np.zeros(2 * col)
    lo.oracle_bmv_f64(m, vp(sy), vp(wt_ref), col, vp(vv), vp(p_ref), vp(info_ref))
    c = np.zeros(4 * m); c[:2 * col] = vv
    cd = _dev(c); wtd2 = _dev(wt_ref.T)
    assert L.lbfgsb_test_dense_f64(3 + warp, m, col, C.c_double(theta), _vp(syd), _vp(wtd2), _vp(cd), vp(info)) == 0
    assert info[0] == info_ref[0] == 0
    assert np.array_equal(cd.cpu().numpy()[2 * m:2 * m + 2 * col], p_ref)
    # formk's tail: WN1 = [Y'ZZ'Y, .; L_a' + R_z', S'AA'S] from a random free/active split
    if warp:
        free = rng.uniform(size=4 * m) < 0.6
        Yf, Sf, Sa, Ya = Y[free], S[free], S[~free], Y[~free]
        wn1 = np.zeros((2 * m, 2 * m), order="F")
        wn1[:col, :col] = np.tril(Yf.T @ Yf)
        wn1[m:m + col, m:m + col] = np.tril(Sa.T @ Sa)
        LR = np.zeros((col, col))
        for i in range(col):
            for j in range(col):
                LR[i, j] = (Sa[:, i] @ Ya[:, j]) if i <= j else (Sf[:, i] @ Yf[:, j])
        wn1[m:m + col, :col] = LR
        wn_ref = np.zeros((2 * m, 2 * m), order="F"); wn1_ref = wn1.copy(order="F")
        lo.oracle_formk_tail_f64(m, col, C.c_double(theta), vp(wn_ref), vp(wn1_ref), vp(sy), vp(info_ref))
        wn1d, wnd = _dev(wn1.T), _dev(np.zeros((2 * m, 2 * m)))
        assert L.lbfgsb_test_dense_f64(15, m, col, C.c_double(theta), _vp(wn1d), _vp(syd), _vp(wnd), vp(info)) == 0
        assert (info[0] != 0) == (info_ref[0] != 0)
        if info_ref[0] == 0:
            got = wnd.cpu().numpy().T
            assert np.array_equal(np.triu(got[:2 * col, :2 * col]), np.triu(wn_ref[:2 * col, :2 * col]))


def test_dpofa_reports_non_positive_pivot(L):
    m = col = 4
    a = np.asfortranarray(np.eye(4))
    a[2, 2] = -1.0
    ref = a.copy(order="F")
    ir = np.zeros(1, np.int32)
    O.lib().oracle_dpofa_f64(ref.ctypes.data_as(C.c_void_p), m, col, ir.ctypes.data_as(C.c_void_p))
    ad = _dev(a.T)
    dummy = _dev(np.zeros(16))
    info = np.zeros(1, np.int32)
    L.lbfgsb_test_dense_f64(0, m, col, C.c_double(1.0), _vp(ad), _vp(dummy), _vp(dummy), info.ctypes.data_as(C.c_void_p))
    assert info[0] == ir[0] == 3


def test_dcsrch_sequence_bitwise(L):
    """dcsrch / dcstep (src/lbfgsb.f90:2942-3415) on a 1-D function with a bracketing phase."""
    def phi(s):
        return -s / (s * s + 2.0), (s * s - 2.0) / (s * s + 2.0) ** 2

    for stp0, stpmax in ((1e-3, 10.0), (10.0, 50.0), (0.1, 1.0)):
        # oracle
        task = H.make_task("START")
        isv = np.zeros(2, np.int32)
        dsv = np.zeros(13)
        stp = np.array([stp0])
        f0, g0 = phi(0.0)
        f = np.array([f0]); g = np.array([g0])
        ref_steps = []
        for _ in range(40):
            O.lib().oracle_dcsrch_f64(f.ctypes.data_as(C.c_void_p), g.ctypes.data_as(C.c_void_p),
                                      stp.ctypes.data_as(C.c_void_p), C.c_double(1e-3), C.c_double(0.9), C.c_double(0.1),
                                      C.c_double(0.0), C.c_double(stpmax), task.ctypes.data_as(C.c_void_p),
                                      isv.ctypes.data_as(C.c_void_p), dsv.ctypes.data_as(C.c_void_p))
            ref_steps.append((H.task_str(task)[:4], stp[0]))
            if H.task_str(task)[:2] != "FG":
                break
            f[0], g[0] = phi(stp[0])
        # device
        code = np.zeros(1, np.int32)   # CS_START
        isv = np.zeros(2, np.int32)
        dsv = np.zeros(13)
        stp = np.array([stp0])
        fv, gv = phi(0.0)
        got = []
        for _ in range(40):
            assert L.lbfgsb_test_dcsrch_f64(C.c_double(fv), C.c_double(gv), stp.ctypes.data_as(C.c_void_p),
                                            C.c_double(stpmax), code.ctypes.data_as(C.c_void_p),
                                            isv.ctypes.data_as(C.c_void_p), dsv.ctypes.data_as(C.c_void_p)) == 0
            name = {1: "FG", 2: "CONV"}.get(int(code[0]), "WARN" if 3 <= code[0] <= 6 else "ERRO")
            got.append((name[:4], stp[0]))
            if code[0] != 1:
                break
            fv, gv = phi(stp[0])
        assert len(got) == len(ref_steps)
        for (a, sa), (b, sb) in zip(got, ref_steps):
            assert a[:2] == b[:2] and sa == sb, (got, ref_steps)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n,m,col,head,frac", [(1000, 5, 5, 1, 0.5), (70001, 10, 7, 4, 0.3), (300000, 20, 20, 13, 0.5),
                                               (4099, 20, 3, 20, 1.0), (50000, 10, 10, 1, 0.001), (33, 7, 6, 2, 0.0)])
def test_formk_delta_matches_the_reference_loops(L, n, m, col, head, frac, dtype):
    """formk's corrections of the old blocks of WN1 for the variables that entered or left the free set
    (src/lbfgsb.f90:1801-1851): the ordered compaction of the listed rows and k_formk_delta on random enter/leave
    sets, against the reference's loops restated here in float64 -- for every (iy, jy), jy <= iy:
    sum over entering rows of wy(k1,ipntr)*wy(k1,jpntr) and ws(k1,ipntr)*ws(k1,jpntr) (:1809-1813), the same over
    leaving rows (:1815-1819); for every (is, jy): ws(k1,ipntr)*wy(k1,jpntr) over each list (:1836-1844).
    The kernel adds the rows in another order, so the gate is a rounding bound: a few hundred eps times the sum of
    the magnitudes of the terms (in REAL64 far below the size of a single term: a missing or doubled row fails)."""
    rng = np.random.default_rng(n + 31 * m + col)
    ldw = (n + 31) // 32 * 32
    ws = rng.standard_normal((m, ldw)).astype(dtype)
    wy = rng.standard_normal((m, ldw)).astype(dtype)
    u = rng.uniform(0.0, 1.0, n)
    state = np.where(u < frac / 2, 1, np.where(u < frac, 2, np.where(u < (1 + frac) / 2, 0, 3))).astype(np.uint8)
    state |= (rng.integers(0, 2, n).astype(np.uint8) << 2)    # bit 2 belongs to another pass: must be ignored
    out = np.zeros(6 * 20 * 20, dtype=dtype)
    wsd, wyd, std = _dev(ws), _dev(wy), _dev(state)
    fn = L.lbfgsb_test_formk_delta_f64 if dtype == np.float64 else L.lbfgsb_test_formk_delta_f32
    fn.argtypes = [C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    assert fn(n, m, col, head, ldw, _vp(wsd), _vp(wyd), _vp(std), out.ctypes.data_as(C.c_void_p)) == 0
    out = out.reshape(6, 20, 20)             # [sum][j][i]: element i + 20 j
    ring = [(head - 1 + i) % m for i in range(col)]
    eps = np.finfo(dtype).eps
    for which, mask in ((0, (state & 3) == 1), (1, (state & 3) == 2)):     # entering, leaving
        Y = wy[ring][:, :n][:, mask].astype(np.float64)
        S = ws[ring][:, :n][:, mask].astype(np.float64)
        for blk, (A, B) in enumerate(((Y, Y), (S, S), (S, Y))):
            ref = A @ B.T                     # ref[i, j] = sum_k A[i, k] B[j, k]
            mag = np.abs(A) @ np.abs(B).T
            got = out[3 * which + blk].T[:col, :col].astype(np.float64)
            for i in range(col):
                for j in range(col if blk == 2 else i + 1):
                    assert abs(got[i, j] - ref[i, j]) <= 600 * eps * mag[i, j] + 1e-300, (which, blk, i, j, got[i, j], ref[i, j])
