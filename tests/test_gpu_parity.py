"""GPU parity tests: the CUDA library (through its C ABI) against the CPU oracle on the same
seeded inputs. Strict-fp contexts must match exactly (bit-exact up to the sign of zero);
fast-fp contexts (FMA contraction, atomic deposit) within the tolerances written below."""
import numpy as np
import pytest

from helpers import (assert_prtls_values_equal, assert_values_equal, random_fields,
                     random_particles, smooth_fields, to_device, to_host)

pytestmark = pytest.mark.gpu

DIMS = {1: (37,), 2: (23, 17), 3: (11, 9, 13)}

# fp32 tolerances of the fast (FMA-contracted) build against the oracle
RTOL_FAST = 2e-5
ATOL_FAST = 2e-6


@pytest.fixture(scope="module")
def eb():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import entity_b200
    return entity_b200


def dev(a):
    import torch
    return torch.from_numpy(a.copy()).cuda()


def host(t):
    return t.cpu().numpy()


@pytest.mark.parametrize("strict", [True, False])
@pytest.mark.parametrize("dim", [1, 2, 3])
@pytest.mark.parametrize("stencil", [None, "ext"])
def test_field_solvers(eb, orc_mod, dim, stencil, strict):
    orc = orc_mod.oracle()
    g = orc_mod.Grid.make(DIMS[dim], 2)
    ctx = eb.Context(DIMS[dim], order=0, strict=strict)
    st = None
    if stencil:
        st = [0.01, 0.02, 0.03, 0.015, 0.012, 0.022, 0.011, 0.017, 0.019]
    em = random_fields(g, 6, 1)
    cur = random_fields(g, 3, 2)
    d_em, d_cur = dev(em), dev(cur)
    orc.faraday(g, em, 0.21, 0.37, st)
    ctx.faraday(d_em, 0.21, 0.37, st)
    orc.ampere(g, em, 0.45, 0.4)
    ctx.ampere(d_em, 0.45, 0.4)
    orc.currents_ampere(g, em, cur, -0.013, 16.0)
    ctx.currents_ampere(d_em, d_cur, -0.013, 16.0)
    if strict:
        assert_values_equal(host(d_em), em, "em")
        assert_values_equal(host(d_cur), cur, "cur")
    else:
        np.testing.assert_allclose(host(d_em), em, rtol=RTOL_FAST, atol=ATOL_FAST)
        np.testing.assert_allclose(host(d_cur), cur, rtol=RTOL_FAST, atol=ATOL_FAST)


@pytest.mark.parametrize("dim", [1, 2, 3])
@pytest.mark.parametrize("fbc", ["periodic", "conductor", "mixed"])
def test_filter_and_ghosts(eb, orc_mod, dim, fbc):
    """CurrentsFilter loop = nfilter x (copy, filter, ghost exchange); strict build, exact."""
    orc = orc_mod.oracle()
    g = orc_mod.Grid.make(DIMS[dim], 2)
    ctx = eb.Context(DIMS[dim], order=0, strict=True)
    P, Cn = orc_mod.FBC_PERIODIC, orc_mod.FBC_CONDUCTOR
    bc = dict(periodic=[P] * 6, conductor=[Cn] * 6, mixed=[P, P, Cn, Cn, P, P])[fbc]
    cur = random_fields(g, 3, 5)
    buff = np.zeros_like(cur)
    d_cur, d_buff = dev(cur), dev(buff)
    # additive sync + ghost fill first, like the engine does after the deposit
    orc.sync_currents(g, cur, buff, bc)
    orc.comm_fields(g, cur, 0, 3, bc)
    ctx.sync_currents(d_cur, d_buff, bc)
    ctx.comm_fields(d_cur, 0, 3, bc)
    assert_values_equal(host(d_cur), cur, "sync+comm")
    nfilter = 3
    for _ in range(nfilter):
        buff[...] = cur
        orc.filter_pass(g, cur, buff, bc)
        orc.comm_fields(g, cur, 0, 3, bc)
    ctx.filter(d_cur, d_buff, nfilter, bc)
    assert_values_equal(host(d_cur), cur, "filter")


PUSH_CASES = [
    dict(pusher_flags=2),
    dict(pusher_flags=4),
    dict(pusher_flags=2 | 8, gca_larmor_max=5.0, gca_e_ovr_b_sqr_max=0.81),
    dict(pusher_flags=1),
    dict(pusher_flags=2, drag_flags=3, sync_coeff=0.01, compton_coeff=0.02),
    dict(pusher_flags=2, pbc="absorb"),
    dict(pusher_flags=4, pbc="reflect"),
    dict(pusher_flags=2, pbc="none", tag_outgoing=1),
    dict(pusher_flags=2, has_atmosphere=1, atm_gx1=-0.3, atm_x_surf=0.4, atm_ds=0.7),
]


def _setup_push(eb, orc_mod, dim, order, case, strict, n=4000):
    ng = orc_mod.nghosts_for(order)
    g = orc_mod.Grid.make(DIMS[dim], ng)
    kw = dict(PUSH_CASES[case])
    pbc = kw.pop("pbc", "periodic")
    code = dict(periodic=orc_mod.PBC_PERIODIC, absorb=orc_mod.PBC_ABSORB,
                reflect=orc_mod.PBC_REFLECT, none=orc_mod.PBC_NONE)[pbc]
    dx = 0.5
    common = dict(dt=0.45 * dx, omegaB0=0.7, mass=1.0, charge=-1.0, dx=dx, xmin=[0.1, 0.2, 0.3],
                  pbc=[code] * 6, **kw)
    octx = orc_mod.make_pusher(**common)
    ctx = eb.Context(DIMS[dim], order=order, strict=strict, dx=dx, xmin=(0.1, 0.2, 0.3))
    gctx = ctx.make_pusher(**common)
    em = smooth_fields(g, 10 + dim, amp=0.6)
    p = random_particles(g, n, 100 + case, umag=2.0, dead_frac=0.05)
    return g, ctx, octx, gctx, em, p, pbc, dx


def _kill_outside(p, g, dim):
    out = np.zeros(p.n, bool)
    for a, nm in enumerate(["i1", "i2", "i3"][:dim]):
        out |= (getattr(p, nm) < 0) | (getattr(p, nm) >= g.n[a])
    p.tag[out] = 0


@pytest.mark.parametrize("dim", [1, 2, 3])
@pytest.mark.parametrize("order", [0, 1, 2, 3])
@pytest.mark.parametrize("case", range(len(PUSH_CASES)))
def test_push_deposit_strict(eb, orc_mod, dim, order, case):
    """Strict build: pusher bit-exact, ordered deposit bit-exact (serial particle order)."""
    orc = orc_mod.oracle()
    g, ctx, octx, gctx, em, p, pbc, dx = _setup_push(eb, orc_mod, dim, order, case, True)
    n = p.n
    d_em = dev(em)
    arr = to_device(p)
    j_ref = np.zeros(g.shape(3), np.float32)
    d_j = dev(j_ref)
    for step in range(3):
        orc.push(g, order, octx, p, n, em)
        ctx.push(gctx, arr, n, d_em)
        q = to_host(arr, n)
        assert_prtls_values_equal(q, p, what=f"push step {step}")
        if pbc == "none":
            _kill_outside(p, g, dim)
            arr = to_device(p)
        orc.deposit(g, order, p, n, -1.0, octx.dt, dx, j_ref)
        ctx.deposit(arr, n, -1.0, octx.dt, d_j, mode=eb.DEPOSIT_ORDERED)
        assert_values_equal(host(d_j), j_ref, f"ordered deposit step {step}")
    assert np.abs(j_ref).max() > 0


@pytest.mark.parametrize("dim", [1, 2, 3])
@pytest.mark.parametrize("order", [0, 1, 2, 3])
@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("mode", ["atomic", "aggregated", "aggregated_sorted"])
def test_push_deposit_fast(eb, orc_mod, dim, order, fused, mode):
    """Fast build (FMA contraction; atomic or warp-aggregated deposit on random-order and on
    cell-sorted particles; optional fusion): fp32 tolerance."""
    orc = orc_mod.oracle()
    g, ctx, octx, gctx, em, p, pbc, dx = _setup_push(eb, orc_mod, dim, order, 0, False, n=20000)
    n = p.n
    if mode == "aggregated_sorted":
        # long runs of same-cell lanes: the case the aggregation is built for
        key = p.i1.astype(np.int64)
        if dim > 1:
            key = key + g.n[0] * p.i2.astype(np.int64)
        if dim > 2:
            key = key + g.n[0] * g.n[1] * p.i3.astype(np.int64)
        perm = np.argsort(key, kind="stable")
        for nm in p.names():
            getattr(p, nm)[:] = getattr(p, nm)[perm]
    dmode = eb.DEPOSIT_ATOMIC if mode == "atomic" else eb.DEPOSIT_AGGREGATED
    d_em = dev(em)
    arr = to_device(p)
    j_ref = np.zeros(g.shape(3), np.float32)
    d_j = dev(j_ref)
    orc.push(g, order, octx, p, n, em)
    orc.deposit(g, order, p, n, -1.0, octx.dt, dx, j_ref)
    if fused:
        ctx.push_deposit(gctx, arr, n, d_em, d_j, mode=dmode)
    else:
        ctx.push(gctx, arr, n, d_em)
        ctx.deposit(arr, n, -1.0, octx.dt, d_j, mode=dmode)
    q = to_host(arr, n)
    for nm in ("i1", "i2", "i3", "i1_prev", "i2_prev", "i3_prev", "tag"):
        # a particle within rounding of a cell face may land on the other side: allow a handful
        assert (getattr(q, nm) != getattr(p, nm)).sum() <= 2, nm
    same = np.ones(n, bool)
    for nm in ("i1", "i2", "i3"):
        same &= getattr(q, nm) == getattr(p, nm)
    for nm in ("ux1", "ux2", "ux3", "dx1", "dx2", "dx3"):
        np.testing.assert_allclose(getattr(q, nm)[same], getattr(p, nm)[same], rtol=1e-4,
                                   atol=2e-5, err_msg=nm)
    scale = np.abs(j_ref).max()
    assert np.abs(host(d_j) - j_ref).max() <= 2e-4 * scale


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_sort_particles(eb, orc_mod, dim):
    g = orc_mod.Grid.make(DIMS[dim], 2)
    ctx = eb.Context(DIMS[dim], order=0)
    n = 5000
    p = random_particles(g, n, 7, dead_frac=0.2)
    arr = to_device(p)
    n_alive = ctx.sort_particles(arr, n, remove_dead=True)
    assert n_alive == int((p.tag == 1).sum())
    q = to_host(arr, n)
    assert (q.tag[:n_alive] == 1).all() and (q.tag[n_alive:] == 0).all()
    key = q.i1.astype(np.int64)
    if dim > 1:
        key = key + g.n[0] * q.i2.astype(np.int64)
    if dim > 2:
        key = key + g.n[0] * g.n[1] * q.i3.astype(np.int64)
    assert (np.diff(key[:n_alive]) >= 0).all(), "not sorted by cell"
    # stable: expected result = numpy stable argsort of the same key
    key0 = p.i1.astype(np.int64)
    if dim > 1:
        key0 = key0 + g.n[0] * p.i2.astype(np.int64)
    if dim > 2:
        key0 = key0 + g.n[0] * g.n[1] * p.i3.astype(np.int64)
    key0[p.tag != 1] = np.iinfo(np.int64).max
    perm = np.argsort(key0, kind="stable")
    for nm in p.names():
        assert np.array_equal(getattr(q, nm), getattr(p, nm)[perm]), nm


def test_sort_payloads_and_skip_prev(eb, orc_mod):
    """Payload planes are permuted with their particle (particles_sort.cpp:150-160, 239-247);
    EB200_SORT_SKIP_PREV permutes everything except i_prev / dx_prev."""
    import torch
    g = orc_mod.Grid.make(DIMS[2], 2)
    ctx = eb.Context(DIMS[2], order=0)
    n = 7000
    p = random_particles(g, n, 11, dead_frac=0.15)
    for flags in (1, 1 | 2):
        arr = to_device(p)
        w = arr["weight"]
        arr["pld_r"] = torch.stack([2.0 * w, w + 1.0, -w]).contiguous()
        arr["pld_i"] = torch.stack([w.view(torch.int32), torch.arange(n, dtype=torch.int32, device="cuda")]).contiguous()
        prev0 = {k: arr[k].clone() for k in ("i1_prev", "i2_prev", "dx1_prev", "dx2_prev")}
        n_alive = ctx.sort_particles(arr, n, remove_dead=flags)
        assert n_alive == int((p.tag == 1).sum())
        key0 = p.i1.astype(np.int64) + g.n[0] * p.i2.astype(np.int64)
        key0[p.tag != 1] = np.iinfo(np.int64).max
        perm = np.argsort(key0, kind="stable")
        assert np.array_equal(arr["pld_i"][1].cpu().numpy(), perm.astype(np.int32))
        w2 = arr["weight"]
        assert torch.equal(arr["pld_r"][0], 2.0 * w2) and torch.equal(arr["pld_r"][1], w2 + 1.0)
        assert torch.equal(arr["pld_r"][2], -w2) and torch.equal(arr["pld_i"][0], w2.view(torch.int32))
        for k, v in prev0.items():
            if flags & 2:
                assert torch.equal(arr[k], v), k  # untouched
            else:
                assert np.array_equal(arr[k].cpu().numpy(), getattr(p, k)[perm]), k


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_sort_particles_counting(eb, orc_mod, dim):
    """EB200_SORT_UNSTABLE: grouped by cell (i1 fastest), not-alive last, the same multiset of
    particles with every array and payload still attached to its particle; the order inside a
    cell is unspecified. Also through the fallback (more cells than twice the particles)."""
    import torch
    g = orc_mod.Grid.make(DIMS[dim], 2)
    for n in (50000, 64):  # 64: fewer particles than cells / 2 -> the radix fallback
        ctx = eb.Context(DIMS[dim], order=0)
        p = random_particles(g, n, 13 + n % 7, dead_frac=0.2)
        arr = to_device(p)
        arr["pld_i"] = torch.arange(n, dtype=torch.int32, device="cuda").reshape(1, -1).contiguous()
        n_alive = ctx.sort_particles(arr, n, remove_dead=1 | 4)
        assert n_alive == int((p.tag == 1).sum())
        q = to_host(arr, n)
        assert (q.tag[:n_alive] == 1).all() and (q.tag[n_alive:] != 1).all()
        key = q.i1.astype(np.int64)
        if dim > 1:
            key = key + g.n[0] * q.i2.astype(np.int64)
        if dim > 2:
            key = key + g.n[0] * g.n[1] * q.i3.astype(np.int64)
        assert (np.diff(key[:n_alive]) >= 0).all(), "not grouped by cell"
        src = arr["pld_i"][0].cpu().numpy()  # where every particle came from
        assert np.array_equal(np.sort(src), np.arange(n))
        for nm in p.names():
            assert np.array_equal(getattr(q, nm), getattr(p, nm)[src]), nm


def test_errors(eb):
    with pytest.raises(eb.EB200Error):
        eb.Context((8, 8), order=12)
    ctx = eb.Context((8, 8), order=0)
    with pytest.raises(eb.EB200Error, match="No particle pusher"):
        ctx.push(ctx.make_pusher(dt=0.1, pusher_flags=0), {}, 0, None)


@pytest.mark.parametrize("which", [1, 2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("strict", [False, True])
@pytest.mark.parametrize("order_kind", ["sorted", "stale", "random"])
def test_fused_kernels_wide_mesh(eb, orc_mod, which, strict, order_kind):
    """Every fused push+deposit kernel (1 per-thread, 2 TMA chunks, 3 four particles per thread,
    4 shared-memory field tile, 5 four per thread gathering from packed nodes, 6 pipelined,
    7 shared-memory resident, 8 moment-accumulating deposit -- fast build only, the strict
    build falls through to kernel 5) on a mesh wide enough for the tile kernel, on cell-sorted
    particles, on a 'stale' order (sorted, then pushed twice without re-sorting: the state
    between two sorts) and on a random order; two consecutive steps so that crossings, periodic
    wraps and strays outside the tile all occur. Strict build: particles bit-exact (the pusher
    arithmetic is the same whichever kernel runs), J within fp32 summation-order tolerance;
    fast build: the tolerances of test_push_deposit_fast."""
    orc = orc_mod.oracle()
    n_cells = (252, 24)  # 252 + 2*2 ghosts = 256 columns: 16-byte aligned rows
    g = orc_mod.Grid.make(n_cells, 2)
    dx = 0.5
    common = dict(dt=0.45 * dx, omegaB0=0.7, mass=1.0, charge=-1.0, dx=dx, xmin=[0.1, 0.2, 0.3],
                  pbc=[orc_mod.PBC_PERIODIC] * 6, pusher_flags=2)
    octx = orc_mod.make_pusher(**common)
    ctx = eb.Context(n_cells, order=0, strict=strict, dx=dx, xmin=(0.1, 0.2, 0.3))
    ctx.set_pd_kernel(which)
    gctx = ctx.make_pusher(**common)
    em = smooth_fields(g, 77, amp=0.6)
    n = 40000 + 3
    p = random_particles(g, n, 4242, umag=1.0, dead_frac=0.03)

    def sort_by_cell(q):
        key = q.i1.astype(np.int64) + g.n[0] * q.i2.astype(np.int64)
        key[q.tag == 0] = 1 << 40
        perm = np.argsort(key, kind="stable")
        for nm in q.names():
            getattr(q, nm)[:] = getattr(q, nm)[perm]

    if order_kind != "random":
        sort_by_cell(p)
    if order_kind == "stale":
        for _ in range(2):
            orc.push(g, 0, octx, p, n, em)
    d_em = dev(em)
    arr = to_device(p)
    launches0 = ctx.launch_count
    for step in range(2):
        j_ref = np.zeros(g.shape(3), np.float32)
        d_j = dev(j_ref)
        orc.push(g, 0, octx, p, n, em)
        orc.deposit(g, 0, p, n, -1.0, octx.dt, dx, j_ref)
        ctx.push_deposit(gctx, arr, n, d_em, d_j, mode=eb.DEPOSIT_AGGREGATED)
        q = to_host(arr, n)
        if strict:
            assert_prtls_values_equal(q, p, what=f"kernel {which} step {step}")
        else:
            same = np.ones(n, bool)
            for nm in ("i1", "i2", "i1_prev", "i2_prev", "tag"):
                same &= getattr(q, nm) == getattr(p, nm)
            assert (~same).sum() <= 4
            for nm in ("ux1", "ux2", "ux3", "dx1", "dx2"):
                np.testing.assert_allclose(getattr(q, nm)[same], getattr(p, nm)[same], rtol=1e-4,
                                           atol=2e-5, err_msg=nm)
            # keep both sides on the same trajectory for the second step
            arr = to_device(p)
        scale = np.abs(j_ref).max()
        assert np.abs(host(d_j) - j_ref).max() <= 2e-4 * scale, f"kernel {which} step {step}"
    assert ctx.launch_count - launches0 >= 2


@pytest.mark.parametrize("strict", [True, False])
@pytest.mark.parametrize("nfilter", [1, 2, 3, 8, 9])
@pytest.mark.parametrize("open_dims", [(1,), (0,), (0, 1)])
@pytest.mark.parametrize("n_cells", [(150, 37), (64, 16), (7, 5)])
def test_fused_filter_open_faces(eb, orc_mod, n_cells, open_dims, nfilter, strict):
    """The fused sweeps on a 2D domain with non-periodic, non-exchanged dimensions (FBC_NONE on
    both faces: the MATCH walls of the reconnection configuration): their ghost cells are static
    during the filter and enter every pass as they are, exactly as in nfilter x (copy,
    DigitalFilter_kernel, ghost fill of the periodic dimensions only). Strict build: identical
    bits, ghost cells included."""
    orc = orc_mod.oracle()
    g = orc_mod.Grid.make(n_cells, 2)
    ctx = eb.Context(n_cells, order=0, strict=strict)
    bc = [orc_mod.FBC_PERIODIC] * 6
    for a in open_dims:
        bc[2 * a] = bc[2 * a + 1] = orc_mod.FBC_NONE
    cur = random_fields(g, 3, 19)  # ghost cells random too: they are the static boundary data
    orc.comm_fields(g, cur, 0, 3, bc)
    d_cur, d_buff = dev(cur), dev(random_fields(g, 3, 23))
    buff = np.zeros_like(cur)
    for _ in range(nfilter):
        buff[...] = cur
        orc.filter_pass(g, cur, buff, bc)
        orc.comm_fields(g, cur, 0, 3, bc)
    ctx.filter(d_cur, d_buff, nfilter, bc)
    if strict:
        assert_values_equal(host(d_cur), cur, f"fused filter, open faces, x{nfilter}")
    else:
        np.testing.assert_allclose(host(d_cur), cur, rtol=RTOL_FAST, atol=ATOL_FAST)


@pytest.mark.parametrize("strict", [True, False])
@pytest.mark.parametrize("nfilter", [1, 2, 4, 5, 8, 9])
@pytest.mark.parametrize("n_cells", [(150, 37), (64, 16), (5, 3)])
def test_fused_filter_passes(eb, orc_mod, n_cells, nfilter, strict):
    """The fused multi-pass filter (temporal blocking, doubly periodic 2D domains) against
    nfilter x (copy, DigitalFilter_kernel, periodic ghost fill) of the oracle: the strict build
    must give identical bits, ghosts included; the fast build within FMA-contraction tolerance.
    Meshes: not a multiple of the tile, exactly one tile, and smaller than the halo (the
    periodic image wraps more than once)."""
    orc = orc_mod.oracle()
    g = orc_mod.Grid.make(n_cells, 2)
    ctx = eb.Context(n_cells, order=0, strict=strict)
    bc = [orc_mod.FBC_PERIODIC] * 6
    cur = random_fields(g, 3, 9)
    orc.comm_fields(g, cur, 0, 3, bc)
    d_cur, d_buff = dev(cur), dev(np.zeros_like(cur))
    buff = np.zeros_like(cur)
    for _ in range(nfilter):
        buff[...] = cur
        orc.filter_pass(g, cur, buff, bc)
        orc.comm_fields(g, cur, 0, 3, bc)
    ctx.filter(d_cur, d_buff, nfilter, bc)
    if strict:
        assert_values_equal(host(d_cur), cur, f"fused filter x{nfilter}")
    else:
        np.testing.assert_allclose(host(d_cur), cur, rtol=RTOL_FAST, atol=ATOL_FAST)


@pytest.mark.parametrize("order_kind", ["sorted", "stale", "random"])
@pytest.mark.parametrize("which", [0, 1, 9])
@pytest.mark.parametrize("weights", ["unit", "spread"])
def test_tile3_kernel(eb, orc_mod, order_kind, which, weights):
    """3D third-order fused push + deposit: the shared-memory fixed-point J tile (kernel 9, what
    0 selects in the fast build) against the oracle, next to the per-lane kernel (1), on
    cell-sorted particles (everything inside the tile), on a stale order (strays take the
    global path) and on a random order (nearly everything does), with unit weights and with
    weights spread over two decades (the tile's scale follows the chunk's largest weight).
    Two consecutive steps; particles to the fast-build tolerances, J to 2e-4 of max|J|."""
    orc = orc_mod.oracle()
    n_cells = (56, 12, 10)
    g = orc_mod.Grid.make(n_cells, orc_mod.nghosts_for(3))
    dx = 0.5
    common = dict(dt=0.45 * dx, omegaB0=0.7, mass=1.0, charge=-1.0, dx=dx, xmin=[0.1, 0.2, 0.3],
                  pbc=[orc_mod.PBC_PERIODIC] * 6, pusher_flags=2)
    octx = orc_mod.make_pusher(**common)
    ctx = eb.Context(n_cells, order=3, strict=False, dx=dx, xmin=(0.1, 0.2, 0.3))
    ctx.set_pd_kernel(which)
    gctx = ctx.make_pusher(**common)
    em = smooth_fields(g, 91, amp=0.6)
    n = 8 * n_cells[0] * n_cells[1] * n_cells[2] + 37
    p = random_particles(g, n, 777, umag=1.0, dead_frac=0.02)
    if weights == "spread":
        rng = np.random.default_rng(5)
        p.weight[:] = (10.0 ** rng.uniform(-1.0, 1.0, n)).astype(np.float32)

    def sort_by_cell(q):
        key = (q.i1.astype(np.int64) + g.n[0] * q.i2.astype(np.int64)
               + g.n[0] * g.n[1] * q.i3.astype(np.int64))
        key[q.tag == 0] = 1 << 40
        perm = np.argsort(key, kind="stable")
        for nm in q.names():
            getattr(q, nm)[:] = getattr(q, nm)[perm]

    if order_kind != "random":
        sort_by_cell(p)
    if order_kind == "stale":
        for _ in range(3):
            orc.push(g, 3, octx, p, n, em)
    d_em = dev(em)
    arr = to_device(p)
    for step in range(2):
        j_ref = np.zeros(g.shape(3), np.float32)
        d_j = dev(j_ref)
        orc.push(g, 3, octx, p, n, em)
        orc.deposit(g, 3, p, n, -1.0, octx.dt, dx, j_ref)
        ctx.push_deposit(gctx, arr, n, d_em, d_j, mode=eb.DEPOSIT_AGGREGATED)
        q = to_host(arr, n)
        same = np.ones(n, bool)
        for nm in ("i1", "i2", "i3", "i1_prev", "i2_prev", "i3_prev", "tag"):
            same &= getattr(q, nm) == getattr(p, nm)
        assert (~same).sum() <= 6
        for nm in ("ux1", "ux2", "ux3", "dx1", "dx2", "dx3"):
            np.testing.assert_allclose(getattr(q, nm)[same], getattr(p, nm)[same], rtol=1e-4,
                                       atol=2e-5, err_msg=nm)
        arr = to_device(p)  # both sides on the same trajectory for the second step
        scale = np.abs(j_ref).max()
        err = np.abs(host(d_j) - j_ref).max()
        assert err <= 2e-4 * scale, f"kernel {which} step {step}: {err / scale:.2e}"
