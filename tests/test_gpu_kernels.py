"""GPU parity tests for the individual sm_100a kernels, all called through the C ABI (ldt_b200.ops -> ctypes).

Checker = oracle/ (CPU restatement pinned to the reference) and, for the NN kernel, the reference's own CUDA kernel
compiled into oracle/_ref/libref_nnd.so.  Bars: bit-exact for distances / indices / the SDE update given identical
inputs; stated tolerances for bf16-input, fp32-accumulate contractions.
"""
import ctypes as C
import os

import pytest
import torch

from ldt_b200 import ops
from oracle import ldt_oracle as O
from tests.helpers import airplane_config, assert_bf16_close, golden, ns, rel_rms_err, rms_rel_err, small_score_cfg

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests selected but no CUDA device"
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def oracle_nn():
    L = C.CDLL(os.path.join(ROOT, "oracle", "liboracle_nn.so"))
    L.oracle_nn_distance.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p] + [C.c_void_p] * 4
    L.oracle_pairwise_cd.argtypes = [C.c_int] * 4 + [C.c_void_p] * 2 + [C.c_int] * 2 + [C.c_void_p, C.c_int]
    return L


def cpu_nn(L, a, b):
    a, b = a.cpu().contiguous(), b.cpu().contiguous()
    bs, n, m = a.shape[0], a.shape[1], b.shape[1]
    d1, d2 = torch.empty(bs, n), torch.empty(bs, m)
    i1, i2 = torch.empty(bs, n, dtype=torch.int32), torch.empty(bs, m, dtype=torch.int32)
    L.oracle_nn_distance(bs, n, a.data_ptr(), m, b.data_ptr(), d1.data_ptr(), i1.data_ptr(), d2.data_ptr(), i2.data_ptr())
    return d1, i1, d2, i2


# ------------------------------------------------------------------------------------------------
# NN distance / Chamfer matrix
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("bs,n,m", [(4, 100, 200), (3, 2048, 2048), (2, 1, 7), (1, 1025, 513), (5, 37, 2049)])
def test_nn_distance_bit_exact_vs_oracle(dev, oracle_nn, bs, n, m):
    from ldt_b200 import ops
    g = torch.Generator().manual_seed(n * 7 + m)
    a, b = torch.rand((bs, n, 3), generator=g), torch.rand((bs, m, 3), generator=g) * 1.3 - 0.1
    d1, i1, d2, i2 = [t.cpu() for t in ops.nn_distance_idx(a.to(dev), b.to(dev))]
    e1, j1, e2, j2 = cpu_nn(oracle_nn, a, b)
    assert torch.equal(i1, j1) and torch.equal(i2, j2)
    assert torch.equal(d1, e1) and torch.equal(d2, e2)  # bit-exact distances


def test_nn_distance_reference_unit_test_and_ties(dev):
    """ChamferDistancePytorch/unit_test.py:22-33 criterion on the golden inputs; duplicated points tie to lowest index."""
    from ldt_b200 import ops
    g = golden("nn.npz")
    d1, i1, d2, i2 = [t.cpu() for t in ops.nn_distance_idx(g["p1"].to(dev), g["p2"].to(dev))]
    assert float(((d1 - g["dist1"]) ** 2).mean() + ((d2 - g["dist2"]) ** 2).mean()) < 1e-8
    assert float((i1 - g["idx1"]).float().norm() + (i2 - g["idx2"]).float().norm()) == 0.0
    d1, i1, d2, i2 = [t.cpu() for t in ops.nn_distance_idx(g["q1"].to(dev), g["q2"].to(dev))]
    assert torch.equal(i1[:, :5], torch.arange(5, dtype=torch.int32).expand(2, 5)) and torch.all(d1[:, :5] == 0)


def test_nn_distance_bit_exact_vs_reference_cuda_kernel(dev):
    """The reference's own NmDistanceKernel (compiled from its source into oracle/_ref) on the same GPU."""
    path = os.path.join(ROOT, "oracle", "_ref", "libref_nnd.so")
    assert os.path.exists(path), "oracle/_ref/libref_nnd.so missing: run `make -C oracle` in the build container"
    R = C.CDLL(path)
    fn = getattr(R, "_Z10nndistanceiiPKfiS0_PfPiS1_S2_P11CUstream_st")  # nndistance(...), nndistance.cu:125
    fn.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p] + [C.c_void_p] * 5
    from ldt_b200 import ops
    for bs, n, m in [(8, 2048, 2048), (3, 700, 1500), (33, 64, 64)]:
        g = torch.Generator().manual_seed(bs + n)
        a = torch.randn((bs, n, 3), generator=g).to(dev)
        b = (torch.randn((bs, m, 3), generator=g) * 0.7).to(dev)
        r1, r2 = torch.empty((bs, n), device=dev), torch.empty((bs, m), device=dev)
        k1 = torch.empty((bs, n), dtype=torch.int32, device=dev)
        k2 = torch.empty((bs, m), dtype=torch.int32, device=dev)
        fn(bs, n, a.data_ptr(), m, b.data_ptr(), r1.data_ptr(), k1.data_ptr(), r2.data_ptr(), k2.data_ptr(),
           torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        d1, i1, d2, i2 = ops.nn_distance_idx(a, b)
        assert torch.equal(i1, k1) and torch.equal(i2, k2)
        assert torch.equal(d1, r1) and torch.equal(d2, r2)


def test_nn_distance_rejects_bad_inputs(dev):
    from ldt_b200 import ops
    a = torch.rand(2, 8, 3, device=dev)
    with pytest.raises(RuntimeError, match="contiguous"):
        ops.nn_distance_idx(a.transpose(0, 1), a)
    with pytest.raises(RuntimeError):
        ops.nn_distance_idx(a.double(), a)
    with pytest.raises(RuntimeError, match="empty"):
        ops.nn_distance_idx(a, torch.rand(2, 0, 3, device=dev))


def test_pairwise_cd_vs_oracle_and_nn_kernel(dev, oracle_nn):
    from ldt_b200 import ops
    g = torch.Generator().manual_seed(5)
    a = torch.randn((7, 300, 3), generator=g)
    b = torch.randn((5, 173, 3), generator=g)
    M = ops.pairwise_cd(a.to(dev), b.to(dev)).cpu()
    ref = torch.empty(7, 5)
    oracle_nn.oracle_pairwise_cd(7, 5, 300, 173, a.data_ptr(), b.data_ptr(), 0, 7, ref.data_ptr(), 4)
    assert torch.allclose(M, ref, rtol=3e-7, atol=0)  # same minima; the sums are both double-accumulated
    # row blocks compose (the multi-GPU sharding unit) and single rows equal the batched NN kernel + means
    blk = ops.pairwise_cd(a.to(dev), b.to(dev), 2, 5).cpu()
    assert torch.equal(blk, M[2:5])
    d1, _, d2, _ = ops.nn_distance_idx(a[3:4].expand(5, -1, -1).contiguous().to(dev), b.to(dev))
    row = (d1.double().mean(1).float() + d2.double().mean(1).float()).cpu()
    assert torch.equal(row, M[3])


def test_pairwise_cd_full_size_properties(dev):
    """BASELINE config-4 point counts (2048 pts): symmetry, zero diagonal, permutation invariance, shard consistency."""
    from ldt_b200 import ops
    g = torch.Generator().manual_seed(7)
    x = torch.randn((24, 2048, 3), generator=g)
    x = x - x.mean(1, keepdim=True)
    x = (x / x.norm(dim=-1).amax(1)[:, None, None]).to(dev)  # unit-sphere normalisation like ShapeNet_55.py:50-54
    M = ops.pairwise_cd(x, x)
    assert torch.equal(M, M.t())
    assert torch.all(M.diagonal() == 0)
    perm = torch.randperm(2048, generator=g).to(dev)
    assert torch.allclose(ops.pairwise_cd(x[:, perm].contiguous(), x), M, rtol=1e-6, atol=0)
    assert torch.equal(torch.cat([ops.pairwise_cd(x, x, 0, 11), ops.pairwise_cd(x, x, 11, 24)]), M)
    assert torch.all(M[~torch.eye(24, dtype=torch.bool, device=dev)] > 0)


def test_cd_metrics_match_reference_golden(dev):
    from ldt_b200 import metrics
    g = golden("metrics.npz")
    res = metrics.compute_CD_metrics(g["smp"].to(dev), g["ref"].to(dev), 4)
    assert float(res["mmd-CD"]) == pytest.approx(float(g["mmd"]), rel=1e-4)
    assert float(res["cov-CD"]) == float(g["cov"])
    assert float(res["1-NN-CD-acc"]) == float(g["acc"])
    M_rs = metrics._pairwise_CD_(g["ref"].to(dev), g["smp"].to(dev), 4).cpu()
    assert torch.allclose(M_rs, g["M_rs"], rtol=1e-4, atol=1e-6)


def _unit_sphere_clouds(n, pts, seed):
    """SURVEY.md 8(d) row 4 inputs: randn clouds centred and scaled to unit max-norm (ShapeNet_55.py:50-54)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((n, pts, 3), generator=g)
    x = x - x.mean(1, keepdim=True)
    return x / x.norm(dim=-1).amax(1)[:, None, None]


def test_pairwise_cd_values_at_2048_points(dev, oracle_nn):
    """The headline CD size value-checked (not only properties): 4 x 4 clouds of 2048 points against (1) the C oracle and
    (2) the reference's own NmDistanceKernel followed by torch's fp32 ``mean(dim=1)`` exactly as _pairwise_CD_ does
    (evaluation_metrics.py:184-191).
    Tolerances: (1) bit-exact -- same minima, both sums double-accumulated and rounded once.  (2) the reference rounds
    every partial sum of its fp32 mean, ours is the correctly rounded mean: <= 4 ulp (rtol 5e-7) is stated; identical
    minima mean cloud-level argmin / topk ties can only flip between entries closer than that."""
    from ldt_b200 import ops
    a, b = _unit_sphere_clouds(4, 2048, 7), _unit_sphere_clouds(4, 2048, 8)
    M = ops.pairwise_cd(a.to(dev), b.to(dev)).cpu()
    ref = torch.empty(4, 4)
    oracle_nn.oracle_pairwise_cd(4, 4, 2048, 2048, a.data_ptr(), b.data_ptr(), 0, 4, ref.data_ptr(), 8)
    assert torch.equal(M, ref)
    path = os.path.join(ROOT, "oracle", "_ref", "libref_nnd.so")
    assert os.path.exists(path), "oracle/_ref/libref_nnd.so missing: run `make -C oracle` in the build container"
    fn = getattr(C.CDLL(path), "_Z10nndistanceiiPKfiS0_PfPiS1_S2_P11CUstream_st")
    fn.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p] + [C.c_void_p] * 5
    ad, bd = a.to(dev), b.to(dev)
    rows = []
    for i in range(4):   # the reference's loop: one row cloud expanded against a batch of column clouds
        ai = ad[i:i + 1].expand(4, -1, -1).contiguous()
        r1, r2 = torch.empty((4, 2048), device=dev), torch.empty((4, 2048), device=dev)
        k1, k2 = torch.empty((4, 2048), dtype=torch.int32, device=dev), torch.empty((4, 2048), dtype=torch.int32, device=dev)
        fn(4, 2048, ai.data_ptr(), 2048, bd.data_ptr(), r1.data_ptr(), k1.data_ptr(), r2.data_ptr(), k2.data_ptr(),
           torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        rows.append((r1.mean(dim=1) + r2.mean(dim=1)).view(1, -1))
    M_ref = torch.cat(rows, dim=0).cpu()
    assert torch.allclose(M, M_ref, rtol=5e-7, atol=0)


def test_pairwise_cd_upper_triangle_mirrors_to_the_full_matrix(dev):
    """The symmetric form (M_rr / M_ss): upper triangle + mirror == the full N^2 evaluation, bit for bit, for every
    interleaved row assignment (the multi-GPU split of SURVEY.md 8e)."""
    from ldt_b200 import metrics, ops
    for n, pts in ((24, 2048), (9, 300), (5, 2500)):
        x = _unit_sphere_clouds(n, pts, 7 + n).to(dev)
        full = ops.pairwise_cd(x, x)
        assert torch.equal(full, full.t())
        up = ops.pairwise_cd_upper(x)
        assert torch.all(torch.tril(up, -1) == 0)                       # nothing below the diagonal is computed
        assert torch.equal(ops.mirror_upper(up), full)
        for world in (2, 3, 8):
            acc = torch.zeros_like(full)
            for r in range(world):
                part = ops.pairwise_cd_upper(x, r, world)
                owned = torch.zeros(n, dtype=torch.bool, device=dev)
                owned[r::world] = True
                assert torch.all(part[~owned] == 0)
                acc += part
            assert torch.equal(ops.mirror_upper(acc), full)
        assert torch.equal(metrics._pairwise_CD_(x, x), full)           # the metrics entry point takes the triangle
    with pytest.raises(RuntimeError, match="row_first"):
        ops.pairwise_cd_upper(x, 3, 2)


def test_compute_cd_metrics_evaluates_two_not_three_matrices(dev):
    """compute_CD_metrics = N^2 (rs) + 2 * N(N+1)/2 (rr, ss upper triangles) pair evaluations; results unchanged."""
    from ldt_b200 import metrics, ops
    ref, smp = _unit_sphere_clouds(12, 256, 3).to(dev), (_unit_sphere_clouds(12, 256, 4) * 0.9).to(dev)
    res = metrics.compute_CD_metrics(smp, ref)
    M_rs, M_rr, M_ss = ops.pairwise_cd(ref, smp), ops.pairwise_cd(ref, ref), ops.pairwise_cd(smp, smp)
    want = {}
    metrics._update(want, metrics.lgan_mmd_cov(M_rs.t()), "CD")
    metrics._update(want, metrics.knn(M_rr, M_rs, M_ss, 1, sqrt=False), "CD", only_acc=True)
    assert set(res) == set(want) and all(torch.equal(res[k], want[k]) for k in want)


# ------------------------------------------------------------------------------------------------
# approximate EMD (ApproxMatch + MatchCost forward)
# ------------------------------------------------------------------------------------------------
def _ref_emd(dev, a, b):
    """The reference's own approxmatch/matchcost kernels (compiled from its source into oracle/_ref) on this GPU."""
    path = os.path.join(ROOT, "oracle", "_ref", "libref_emd.so")
    assert os.path.exists(path), "oracle/_ref/libref_emd.so missing: run `make -C oracle` in the build container"
    R = C.CDLL(path)
    am = getattr(R, "_Z11approxmatchiiiPKfS0_PfS1_P11CUstream_st")   # approxmatch(...), approxmatch.cu:299
    mc = getattr(R, "_Z9matchcostiiiPKfS0_PfS1_P11CUstream_st")      # matchcost(...),   approxmatch.cu:309
    am.argtypes = [C.c_int] * 3 + [C.c_void_p] * 5
    mc.argtypes = [C.c_int] * 3 + [C.c_void_p] * 5
    bs, n, m = a.shape[0], a.shape[1], b.shape[1]
    match = torch.zeros((bs, m, n), device=dev)
    temp = torch.zeros((bs, (n + m) * 2), device=dev)
    cost = torch.zeros((bs,), device=dev)
    torch.cuda.synchronize()
    am(bs, n, m, a.data_ptr(), b.data_ptr(), match.data_ptr(), temp.data_ptr(), None)
    mc(bs, n, m, a.data_ptr(), b.data_ptr(), match.data_ptr(), cost.data_ptr(), None)
    torch.cuda.synchronize()
    return match, cost


@pytest.mark.parametrize("bs,n,m", [(3, 256, 256), (2, 2048, 2048), (33, 100, 100), (2, 512, 256), (2, 300, 900), (1, 1, 1)])
def test_match_cost_vs_reference_cuda_kernels(dev, bs, n, m):
    """Fused match cost (no dense match) and the materialised match against the reference kernels on the same inputs.
    Tolerance: the fused kernel sums w*d per level instead of (sum of w)*d per entry -> fp32 association differences."""
    from ldt_b200 import ops
    g = torch.Generator().manual_seed(bs * 1000 + n + m)
    a = (torch.rand((bs, n, 3), generator=g) - 0.5).to(dev)
    b = (torch.rand((bs, m, 3), generator=g) - 0.5).to(dev) * 0.9
    ref_match, ref_cost = _ref_emd(dev, a, b)
    cost = ops.match_cost(a, b)
    assert torch.allclose(cost, ref_cost, rtol=5e-5, atol=1e-6), (cost, ref_cost)
    match = ops.approx_match(a, b)
    assert torch.allclose(match, ref_match, rtol=1e-4, atol=1e-7), (match - ref_match).abs().max()
    cost2 = ops.match_cost_from_match(a, b, match)
    assert torch.allclose(cost2, ref_cost, rtol=5e-5, atol=1e-6)


def test_match_cost_vs_cpu_oracle_and_pairwise_matrix(dev):
    from ldt_b200 import metrics, ops
    L = C.CDLL(os.path.join(ROOT, "oracle", "liboracle_emd.so"))
    L.oracle_match_cost.argtypes = [C.c_int] * 3 + [C.c_void_p] * 4
    g = torch.Generator().manual_seed(3)
    a = torch.randn((4, 128, 3), generator=g) * 0.3
    b = torch.randn((4, 128, 3), generator=g) * 0.3
    want = torch.empty(4)
    L.oracle_match_cost(4, 128, 128, a.data_ptr(), b.data_ptr(), want.data_ptr(), None)
    got = ops.match_cost(a.to(dev), b.to(dev)).cpu()
    assert torch.allclose(got, want, rtol=2e-4, atol=1e-6), (got, want)   # CPU expf vs GPU ex2.approx
    # matrix form == per-pair match cost / points, rows sharded or not
    M = ops.pairwise_emd(a.to(dev), b.to(dev))
    for i in range(4):
        row = ops.match_cost(a[i:i + 1].expand(4, -1, -1).contiguous().to(dev), b.to(dev)) / 128.0
        assert torch.equal(M[i], row)
    assert torch.equal(ops.pairwise_emd(a.to(dev), b.to(dev), 1, 3), M[1:3])
    # identical sets: every cloud is its own nearest neighbour under EMD
    S = ops.pairwise_emd(a.to(dev), a.to(dev))
    assert torch.equal(S.argmin(dim=1).cpu(), torch.arange(4))
    res = metrics.compute_all_metrics(a.to(dev), b.to(dev), batch_size=2)
    assert set(res) == {"mmd-CD", "cov-CD", "mmd-EMD", "cov-EMD", "1-NN-CD-acc", "1-NN-EMD-acc"}


# ------------------------------------------------------------------------------------------------
# GEMM core
# ------------------------------------------------------------------------------------------------
def _gemm_ref(A, W, bias, epi, resid=None, gate=None, rows_per_gate=1):
    acc = A.float() @ W.float().t() + bias
    if epi == 2:
        acc = torch.nn.functional.gelu(acc)
    if epi == 3:
        g = 1.0 if gate is None else gate.repeat_interleave(rows_per_gate, dim=0)[: acc.shape[0]]
        acc = resid + g * acc
    return acc


@pytest.mark.parametrize("backend", [0, 1, 2, 3])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 128, 256), (256, 512, 1024), (64, 120, 1024), (200, 1024, 128),
                                   (4096, 3072, 1024), (1000, 149504 // 8, 1024), (96, 8, 128)])
def test_gemm_all_epilogues(dev, backend, M, N, K):
    from ldt_b200 import ops
    if backend == 1 and M * N * K > 2 ** 31:
        pytest.skip("naive cross-check kernel: skip the largest shapes")
    g = torch.Generator().manual_seed(M + N + K)
    A = (torch.randn((M, K), generator=g) * 0.5).to(dev).bfloat16()
    W = (torch.randn((N, K), generator=g) / K ** 0.5).to(dev).bfloat16()
    bias = torch.randn((N,), generator=g).to(dev)
    ldo = (N + 7) // 8 * 8
    for epi, dt in ((0, torch.float32), (1, torch.bfloat16), (2, torch.bfloat16), (3, torch.float32)):
        out = torch.full((M, ldo), 7.0, dtype=dt, device=dev)
        kw = {}
        resid = gate = None
        rpg = 32
        if epi == 3:
            resid = torch.randn((M, ldo), generator=g).to(dev)
            gate = torch.randn(((M + rpg - 1) // rpg, N), generator=g).to(dev)
            out = resid.clone()
            kw = dict(resid=out, gate=gate, gate_stride=N, rows_per_gate=rpg)
        ops.gemm(A, W, bias, out, epi, N=N, K=K, backend=backend, **kw)
        ref = _gemm_ref(A, W, bias, epi, None if resid is None else resid[:, :N], gate, rpg)
        got = out[:, :N].float()
        if dt == torch.bfloat16:
            assert_bf16_close(got, ref, 1e-4, f"backend {backend} epi {epi}")
        else:  # fp32 output of exact bf16 products accumulated in fp32: only summation order differs
            err = rel_rms_err(got, ref)
            assert err < 1e-4, f"backend {backend} epi {epi}: rel err {err}"
        if ldo > N and epi != 3:
            assert torch.all(out[:, N:].float() == 7.0), "wrote outside the N columns"


@pytest.mark.parametrize("M,N,K", [(256, 256, 32), (512, 1024, 1024), (8192, 1024, 4096), (300, 120, 128), (64, 3072, 1024),
                                   (2048, 4096, 1024)])
def test_gemm_tf32_operands(dev, M, N, K):
    """operand_type 1: fp32 operands through the CTA-pair pipeline as kind::tf32 (the TF32 parity mode).  (1) small-integer
    inputs: exact, bit-equal to fp64 -- catches tensor-map / descriptor / K-advance mistakes of the f32 layout; (2) TF32-
    rounded random inputs: every product is exact in fp32, only the summation order differs from torch's fp32 matmul
    (run with allow_tf32 off); (3) the three epilogues available to f32 operands."""
    from ldt_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randint(-4, 5, (M, K), generator=g).float().to(dev)
    W = torch.randint(-4, 5, (N, K), generator=g).float().to(dev)
    bias = torch.randint(-8, 9, (N,), generator=g).float().to(dev)
    ldo = (N + 7) // 8 * 8
    out = torch.full((M, ldo), 7.0, device=dev)
    ops.gemm(A, W, bias, out, 0, N=N, K=K)
    assert torch.equal(out[:, :N], (A.double() @ W.double().t() + bias.double()).float())
    if ldo > N:
        assert torch.all(out[:, N:] == 7.0)
    A = O.tf32_round(torch.randn((M, K), generator=g) * 0.5).to(dev)
    W = O.tf32_round(torch.randn((N, K), generator=g) / K ** 0.5).to(dev)
    bias = torch.randn((N,), generator=g).to(dev)
    acc = (A.double() @ W.double().t() + bias.double())
    out = torch.empty((M, ldo), device=dev)
    ops.gemm(A, W, bias, out, 0, N=N, K=K)
    assert rel_rms_err(out[:, :N], acc) < 5e-5, rel_rms_err(out[:, :N], acc)   # tensor-core fp32 accumulation order / alignment
    ops.gemm(A, W, bias, out, 4, N=N, K=K)     # tf32_round(gelu_erf(acc + bias))
    ref = O.tf32_round(torch.nn.functional.gelu(acc.float().cpu())).to(dev)
    # one TF32 ulp (2^-10 relative) where the fp32 sum lands on the other side of a rounding boundary, plus the fp32
    # accumulation noise of the pre-activation
    assert torch.allclose(out[:, :N], ref, rtol=2.0 ** -9, atol=2e-5), (out[:, :N] - ref).abs().max()
    rpg = 32
    resid = torch.randn((M, ldo), generator=g).to(dev)
    gate = torch.randn(((M + rpg - 1) // rpg, N), generator=g).to(dev)
    out = resid.clone()
    ops.gemm(A, W, bias, out, 3, N=N, K=K, resid=out, gate=gate, gate_stride=N, rows_per_gate=rpg)
    ref = resid[:, :N].double() + gate.double().repeat_interleave(rpg, 0)[:M] * acc
    assert rel_rms_err(out[:, :N], ref) < 2e-4   # worst element over 8 M outputs; gate * acc scales the accumulation noise
    with pytest.raises(RuntimeError):
        ops.gemm(A, W, bias, torch.empty((M, ldo), dtype=torch.bfloat16, device=dev), 1, N=N, K=K)   # no bf16 epilogue for f32 operands


@pytest.mark.parametrize("M,N,K", [(300, 128, 259), (4096, 128, 128), (64, 256, 3), (2, 256, 256), (8192, 1024, 128)])
def test_split_tf32_contraction_is_fp32_grade_with_relu_epilogues(dev, M, N, K):
    """"3xTF32": ldt_split_tf32 lays activations out as [hi | hi | lo] and weights as [hi | lo | hi]; one kind::tf32
    contraction over 3 K then reproduces the fp32 product to ~1e-5 (plain TF32: ~8e-4), which is what lets the fp32
    Conv1d / Linear layers of the encoder and condition prologues run on the tensor cores.  Also the two ReLU epilogues
    (Conv + BatchNorm + ReLU, ConvBNReLURes1D's residual form) and the bit pattern of the split itself."""
    g = torch.Generator().manual_seed(M * 7 + N + K)
    A = (torch.randn((M, K), generator=g) * 0.7).to(dev)
    W = (torch.randn((N, K), generator=g) / K ** 0.5).to(dev)
    bias = torch.randn((N,), generator=g).to(dev)
    ld = (K + 31) // 32 * 32
    A3, W3 = ops.split_tf32(A), ops.split_tf32(W, weight_side=True)
    assert A3.shape == (M, 3 * ld) and W3.shape == (N, 3 * ld)
    hi = O.tf32_round(A.cpu())
    lo = O.tf32_round(A.cpu() - hi)
    assert torch.equal(A3[:, :K].cpu(), hi) and torch.equal(A3[:, ld:ld + K].cpu(), hi) and torch.equal(A3[:, 2 * ld:2 * ld + K].cpu(), lo)
    assert torch.all(A3[:, K:ld] == 0) and torch.all(A3[:, 2 * ld + K:] == 0)
    whi = O.tf32_round(W.cpu())
    assert torch.equal(W3[:, ld:ld + K].cpu(), O.tf32_round(W.cpu() - whi)) and torch.equal(W3[:, 2 * ld:2 * ld + K].cpu(), whi)
    acc = A.double() @ W.double().t() + bias.double()
    out = torch.empty((M, N), device=dev)
    ops.gemm(A3, W3, bias, out, 0)
    e3 = rel_rms_err(out, acc)
    ops.gemm(ops.round_pad_tf32(A, ld), ops.round_pad_tf32(W, ld), bias, out, 0)
    e1 = rel_rms_err(out, acc)
    # the floor is the tensor core's fp32 accumulation (~7e-6 on exact products, test_gemm_tf32_operands), not the operands
    assert e3 < 2e-5 and (K < 32 or e3 < e1 / 30), (e3, e1)
    ops.gemm(A3, W3, bias, out, 5)                                   # LDT_EPI_BIAS_RELU_F32
    assert torch.all(out >= 0) and rel_rms_err(out, acc.clamp_min(0)) < 2e-5
    resid = torch.randn((M, N), generator=g).to(dev)
    out = resid.clone()
    ops.gemm(A3, W3, bias, out, 6, resid=out)                        # LDT_EPI_RESID_RELU_F32, in place
    assert torch.all(out >= 0) and rel_rms_err(out, (acc + resid.double()).clamp_min(0)) < 2e-5
    with pytest.raises(RuntimeError):
        ops.gemm(A3, W3, bias, out, 6, resid=out, gate=bias)         # the ReLU residual form takes no gate


@pytest.mark.parametrize("normalize", ["anchor", "center", None])
def test_group_features_and_group_max_vs_torch_expression(dev, normalize):
    """ldt_group_features == LocalGrouper's gather / normalise / affine / concatenate (model/Compressor/layers.py:300-317)
    written out in torch on the same indices; ldt_group_max == amax over the neighbours."""
    g = torch.Generator().manual_seed(5)
    B, N, D, S, k = 3, 500, 128, 32, 24
    xyz = torch.randn((B, N, 3), generator=g).to(dev)
    fea = torch.randn((B, N, D), generator=g).to(dev)
    ci = torch.stack([torch.randperm(N, generator=g)[:S] for _ in range(B)]).int().to(dev)
    gi = torch.randint(0, N, (B, S, k), generator=g).int().to(dev)
    alpha = (torch.rand((1, 1, 1, D + 3), generator=g) + 0.5).to(dev)
    beta = torch.randn((1, 1, 1, D + 3), generator=g).to(dev)
    rows = ops.group_features(xyz, fea, ci, gi, normalize, alpha, beta)
    ld = rows.shape[1]
    assert ld == 288 and rows.shape[0] == B * S * k

    def take(p, idx):
        flat = idx.reshape(B, -1).long()
        return torch.gather(p, 1, flat.unsqueeze(-1).expand(-1, -1, p.shape[-1])).reshape(*idx.shape, p.shape[-1])
    anchor = take(fea, ci)
    grouped = torch.cat([take(fea, gi), take(xyz, gi)], dim=-1)
    if normalize is not None:
        mean = grouped.mean(dim=2, keepdim=True) if normalize == "center" else torch.cat([anchor, take(xyz, ci)], dim=-1).unsqueeze(-2)
        centred = grouped - mean
        std = torch.std(centred.reshape(B, -1), dim=-1, keepdim=True)[:, :, None, None]
        grouped = alpha * (centred / (std + 1e-5)) + beta
    ref = torch.cat([grouped, anchor.unsqueeze(2).expand(-1, -1, k, -1)], dim=-1).reshape(B * S * k, 2 * D + 3)
    assert torch.all(rows[:, 2 * D + 3:] == 0)
    assert torch.allclose(rows[:, :2 * D + 3], ref, rtol=2e-5, atol=2e-6), (rows[:, :2 * D + 3] - ref).abs().max()
    if normalize is None:
        assert torch.equal(rows[:, :2 * D + 3], ref)                 # pure gather + concatenate
    again = ops.group_features(xyz, fea, ci, gi, normalize, alpha, beta)
    assert torch.equal(rows, again)                                  # fixed-order sums: deterministic
    mx = ops.group_max(rows, k)
    assert torch.equal(mx, rows.reshape(B * S, k, ld).amax(dim=1))


def test_gemm_tcgen05_matches_cross_check_bitwise_on_exact_inputs(dev):
    """Small-integer inputs make every product and partial sum exact, so tcgen05, the SIMT cross-check and fp64
    must agree to the bit: catches descriptor / swizzle / K-advance mistakes that tolerance tests can hide."""
    from ldt_b200 import ops
    g = torch.Generator().manual_seed(3)
    for M, N, K in [(128, 256, 64), (384, 256, 512), (130, 136, 192), (8192, 1024, 1024)]:
        A = torch.randint(-4, 5, (M, K), generator=g).float().to(dev).bfloat16()
        W = torch.randint(-4, 5, (N, K), generator=g).float().to(dev).bfloat16()
        bias = torch.randint(-8, 9, (N,), generator=g).float().to(dev)
        ref = (A.double() @ W.double().t() + bias.double()).float()
        for backend in (0, 2, 3):  # library's choice, single-CTA tiles, CTA-pair (cta_group::2) tiles
            o0 = torch.empty((M, N), device=dev)
            ops.gemm(A, W, bias, o0, 0, backend=backend)
            assert torch.equal(o0, ref), f"tcgen05 GEMM (backend {backend}) wrong at {M}x{N}x{K}: max diff {(o0 - ref).abs().max()}"
        if M * N * K <= 2 ** 28:
            o1 = torch.empty((M, N), device=dev)
            ops.gemm(A, W, bias, o1, 0, backend=1)
            assert torch.equal(o1, ref)


@pytest.mark.parametrize("M,N,K", [(512, 256, 64), (768, 256, 256), (2304, 1024, 4096), (8192, 1024, 1024), (1000, 512, 512)])
def test_gemm_cluster4_multicast_equals_pair_kernel_bitwise(dev, M, N, K):
    """backend 4 = clusters of two CTA pairs sharing the W tile by TMA multicast (gemm_tc4_kernel): same tile arithmetic
    as the pair kernel, so every epilogue must give the same bits, including a ragged / odd number of 256-row tiles."""
    from ldt_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    A = (torch.randn((M, K), generator=g) * 0.5).to(dev).bfloat16()
    W = (torch.randn((N, K), generator=g) / K ** 0.5).to(dev).bfloat16()
    bias = torch.randn((N,), generator=g).to(dev)
    resid = torch.randn((M, N), generator=g).to(dev)
    gate = torch.randn(((M + 31) // 32, N), generator=g).to(dev)
    for epi, dt in ((0, torch.float32), (1, torch.bfloat16), (2, torch.bfloat16), (3, torch.float32)):
        outs = []
        for backend in (3, 4):
            out = resid.clone() if epi == 3 else torch.full((M, N), 7.0, dtype=dt, device=dev)
            kw = dict(resid=out, gate=gate, gate_stride=N, rows_per_gate=32) if epi == 3 else {}
            ops.gemm(A, W, bias, out, epi, backend=backend, **kw)
            outs.append(out)
        assert torch.equal(outs[0], outs[1]), f"epilogue {epi}: max diff {(outs[0].float() - outs[1].float()).abs().max()}"


@pytest.mark.parametrize("M,Cc,inner", [(8192, 1024, 4096), (2048, 1024, 4096), (1056, 256, 1024), (4096, 512, 512), (1024, 256, 256)])
def test_fused_mlp_equals_two_gemm_launches_bitwise(dev, M, Cc, inner):
    """ldt_mlp_bf16 (fc1 + GELU -> fc2 + gate + residual in one persistent kernel, fc2 tiles released per 256-row block
    by global completion counters) runs the same tile arithmetic as the two ldt_gemm_bf16 launches: bit-identical
    outputs AND hidden activations; the counters are left zero, so the same buffer serves the next launch."""
    from ldt_b200 import ops
    g = torch.Generator().manual_seed(M + Cc + inner)
    A = (torch.randn((M, Cc), generator=g) * 0.7).to(dev).bfloat16()
    W1 = (torch.randn((inner, Cc), generator=g) / Cc ** 0.5).to(dev).bfloat16()
    W2 = (torch.randn((Cc, inner), generator=g) / inner ** 0.5).to(dev).bfloat16()
    b1 = torch.randn((inner,), generator=g).to(dev)
    b2 = torch.randn((Cc,), generator=g).to(dev)
    resid = torch.randn((M, Cc), generator=g).to(dev)
    rpg = 32
    gate = torch.randn(((M + rpg - 1) // rpg, Cc), generator=g).to(dev)
    # the two-launch path
    hid_ref = torch.empty((M, inner), dtype=torch.bfloat16, device=dev)
    out_ref = resid.clone()
    ops.gemm(A, W1, b1, hid_ref, 2, backend=3)
    ops.gemm(hid_ref, W2, b2, out_ref, 3, resid=out_ref, gate=gate, gate_stride=Cc, rows_per_gate=rpg, backend=3)
    # against plain torch (same bf16 rounding point for the hidden activations)
    h32 = torch.nn.functional.gelu(A.float() @ W1.float().t() + b1).bfloat16().float()
    ref = resid + gate.repeat_interleave(rpg, 0)[:M] * (h32 @ W2.float().t() + b2)
    assert rel_rms_err(out_ref, ref) < 4e-3
    sync = ops.mlp_sync_buffer(M, dev)
    for rep in range(3):   # repeated launches on one counter buffer: the kernel must leave it clean
        hid = torch.zeros((M, inner), dtype=torch.bfloat16, device=dev)
        out = resid.clone()
        ops.mlp(A, W1, b1, hid, W2, b2, out, sync, resid=out, gate=gate, gate_stride=Cc, rows_per_gate=rpg)
        torch.cuda.synchronize()
        assert torch.equal(hid, hid_ref), f"hidden activations differ (launch {rep})"
        assert torch.equal(out, out_ref), f"fused MLP output differs (launch {rep}): max {(out - out_ref).abs().max()}"
        assert int(sync.abs().sum()) == 0, "completion counters not left clean"
    # broadcast gate row (unconditional sampling: one AdaLN row for the whole batch) and no gate at all
    out = resid.clone()
    ops.mlp(A, W1, b1, hid, W2, b2, out, sync, resid=out, gate=gate[:1], gate_stride=0, rows_per_gate=rpg)
    o2 = resid.clone()
    ops.gemm(hid_ref, W2, b2, o2, 3, resid=o2, gate=gate[:1], gate_stride=0, rows_per_gate=rpg, backend=3)
    assert torch.equal(out, o2)


def test_fused_mlp_rejects_unsupported_shapes(dev):
    from ldt_b200 import ops
    A = torch.zeros((1024, 128), dtype=torch.bfloat16, device=dev)
    W1 = torch.zeros((512, 128), dtype=torch.bfloat16, device=dev)
    W2 = torch.zeros((128, 512), dtype=torch.bfloat16, device=dev)
    hid = torch.zeros((1024, 512), dtype=torch.bfloat16, device=dev)
    out = torch.zeros((1024, 128), device=dev)
    with pytest.raises(RuntimeError, match="multiples of 256"):
        ops.mlp(A, W1, None, hid, W2, None, out, ops.mlp_sync_buffer(1024, dev))


# ------------------------------------------------------------------------------------------------
# element-wise kernels
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,C_", [(64, 128), (96, 1024), (8192, 1024), (33, 256)])
def test_layernorm_modulate(dev, rows, C_):
    from ldt_b200 import ops
    g = torch.Generator().manual_seed(rows + C_)
    x = (torch.randn((rows, C_), generator=g) * 3 + 1).to(dev)
    nb = (rows + 31) // 32
    mod = torch.randn((nb, 3 * C_), generator=g).to(dev)
    y = torch.empty((rows, C_), dtype=torch.bfloat16, device=dev)
    ops.layernorm_mod(x, y, shift=mod[:, :C_], scale=mod[:, C_:2 * C_], mod_stride=3 * C_, rows_per_mod=32)
    n = torch.nn.functional.layer_norm(x, (C_,), eps=1e-6)
    sh = mod[:, :C_].repeat_interleave(32, 0)[:rows]
    sc = mod[:, C_:2 * C_].repeat_interleave(32, 0)[:rows]
    ref = n * (1 + sc) + sh  # modulate(), model/layers.py:136-137
    assert_bf16_close(y.float(), ref, 1e-5, "adaLN")
    # broadcast (stride 0) AdaLN row and the affine flavour used by the decoder blocks
    ops.layernorm_mod(x, y, shift=mod[:1, :C_], scale=mod[:1, C_:2 * C_], mod_stride=0, rows_per_mod=32)
    assert_bf16_close(y.float(), n * (1 + mod[0, C_:2 * C_]) + mod[0, :C_], 1e-5, "adaLN broadcast")
    w, b = mod[0, :C_].contiguous(), mod[0, C_:2 * C_].contiguous()
    ops.layernorm_mod(x, y, weight=w, bias=b)
    assert_bf16_close(y.float(), torch.nn.functional.layer_norm(x, (C_,), w, b, 1e-6), 1e-5, "affine")


def test_cast_pad(dev):
    from ldt_b200 import ops
    x = torch.randn(96, 120, device=dev)
    y = ops.cast_pad_bf16(x, 128)
    assert torch.equal(y[:, :120], x.bfloat16()) and torch.all(y[:, 120:] == 0)
    w = ops.pack_weight(torch.randn(40, 20, 1, device=dev))
    assert w.shape == (40, 64) and torch.all(w[:, 20:] == 0)


def test_time_embedding_vs_oracle(dev):
    from ldt_b200 import ops
    D, half = 256, 32
    g = torch.Generator().manual_seed(2)
    sd = O.synth_state_dict({"TimeEmbedding.mlp.0.weight": (D, 2 * half), "TimeEmbedding.mlp.0.bias": (D,),
                             "TimeEmbedding.mlp.2.weight": (D, D), "TimeEmbedding.mlp.2.bias": (D,)}, 4)
    for R in (1, 5, 16, 300):
        t = torch.rand((R,), generator=g) * (1 - 1e-6) + 1e-6
        ref = O.time_embedding(sd, t)
        c = torch.empty((R, D), device=dev)
        sc = torch.empty((R, D), dtype=torch.bfloat16, device=dev)
        scratch = torch.empty((R, D + 2 * half), device=dev)
        ops.time_embedding(t.to(dev), O.time_freq(2 * half).to(dev), sd["TimeEmbedding.mlp.0.weight"].to(dev),
                           sd["TimeEmbedding.mlp.0.bias"].to(dev), sd["TimeEmbedding.mlp.2.weight"].to(dev),
                           sd["TimeEmbedding.mlp.2.bias"].to(dev), None, c, sc, scratch)
        assert rel_rms_err(c, ref) < 1e-5, rel_rms_err(c, ref)
        assert_bf16_close(sc.float(), torch.nn.functional.silu(ref), 1e-5, "silu(c)")
    extra = torch.randn((300, D), generator=g)
    ops.time_embedding(t.to(dev), O.time_freq(2 * half).to(dev), sd["TimeEmbedding.mlp.0.weight"].to(dev),
                       sd["TimeEmbedding.mlp.0.bias"].to(dev), sd["TimeEmbedding.mlp.2.weight"].to(dev),
                       sd["TimeEmbedding.mlp.2.bias"].to(dev), extra.to(dev), c, sc, scratch)
    assert rel_rms_err(c, ref + extra) < 1e-5


@pytest.mark.parametrize("B,H,Nq,dh", [(3, 16, 32, 64), (2, 4, 2048, 32), (2, 4, 1000, 32), (5, 2, 32, 64),
                                       (3, 16, 32, 8), (2, 8, 32, 16), (2, 16, 100, 8)])
def test_attention_with_reference_layout_quirk(dev, B, H, Nq, dh):
    from ldt_b200 import ops
    C_ = H * dh
    g = torch.Generator().manual_seed(B * Nq)
    q = (torch.randn((B * Nq, C_), generator=g)).to(dev).bfloat16()
    kv = (torch.randn((B * 32, 2 * C_), generator=g)).to(dev).bfloat16()
    o = torch.empty((B * Nq, C_), dtype=torch.bfloat16, device=dev)
    from ldt_b200.score import _PtrView
    ops.attention_nk32(B, H, Nq, dh, q, C_, kv, _PtrView(kv.data_ptr() + 2 * C_), 2 * C_, o)
    # oracle semantics (model/layers.py:192-197) on the same bf16-rounded operands, channels-first
    qf = q.float().view(B, Nq, C_).transpose(1, 2)
    kf = kv.float()[:, :C_].reshape(B, 32, C_).transpose(1, 2)
    vf = kv.float()[:, C_:].reshape(B, 32, C_).transpose(1, 2)
    qh = qf.reshape(B, H, dh, Nq).permute(0, 1, 3, 2)
    kh = kf.reshape(B, H, dh, 32).permute(0, 1, 3, 2)
    vh = vf.reshape(B, H, dh, 32).permute(0, 1, 3, 2)
    w = ((qh @ kh.transpose(-2, -1)) * dh ** -0.5).softmax(-1)
    ref = (w @ vh).reshape(B, Nq, C_)  # == the token-major buffer the next layer reads (quirk: no head permute)
    # P is rounded to bf16 before P.V and the output once more: rms error ~2^-9, worst element a few times that
    assert rms_rel_err(o.float().view(B, Nq, C_), ref) < 4e-3
    assert rel_rms_err(o.float().view(B, Nq, C_), ref) < 3e-2


@pytest.mark.parametrize("B,H,Nq,dh", [(3, 16, 32, 64), (9, 16, 32, 64), (2, 4, 2048, 32), (2, 4, 1000, 32), (1, 2, 128, 64),
                                       (6, 4, 32, 32), (64, 16, 32, 64)])
def test_tcgen05_attention_equals_mma_sync_attention(dev, B, H, Nq, dh):
    """csrc/attention_tc.cu (S = Q K^T and O = P V as tcgen05.mma, block-diagonal P for four stacked samples) against the
    warp-level mma.sync kernel on the same operands.  Both round exp2((s - max) * scale) to bf16 before P V and divide by
    the fp32 row sum afterwards; what differs is the fp32 summation order inside the tensor core, i.e. a rare 1-ulp flip
    of a bf16 probability or output."""
    from ldt_b200 import ops
    from ldt_b200.score import _PtrView
    C_ = H * dh
    g = torch.Generator().manual_seed(B * Nq + dh)
    q = (torch.randn((B * Nq, C_), generator=g)).to(dev).bfloat16()
    kv = (torch.randn((B * 32, 2 * C_), generator=g)).to(dev).bfloat16()
    outs = []
    try:
        for backend in (1, 0):
            ops.set_attention_backend(backend)
            o = torch.full((B * Nq, C_), 7.0, dtype=torch.bfloat16, device=dev)
            ops.attention_nk32(B, H, Nq, dh, q, C_, kv, _PtrView(kv.data_ptr() + 2 * C_), 2 * C_, o)
            torch.cuda.synchronize()
            outs.append(o.float())
    finally:
        ops.set_attention_backend(0)
    sync, tc = outs
    assert torch.isfinite(tc).all()
    assert rms_rel_err(tc, sync) < 1.5e-3, rms_rel_err(tc, sync)
    assert rel_rms_err(tc, sync) < 2e-2, rel_rms_err(tc, sync)


@pytest.mark.parametrize("B", [1, 8, 13, 256])
def test_fused_qkv_attention_equals_unfused_path(dev, B):
    """ldt_qkv_attention_bf16 (head-major packed weights, attention in the GEMM epilogue) against the unfused
    GEMM -> [M,3072] bf16 -> attention kernel pipeline on the same operands: same rounding points (q,k,v and the
    softmax numerators rounded to bf16), so the two must agree to accumulation-order noise; and both against the
    fp32 formula of model/layers.py:186-197 on the bf16-rounded operands."""
    from ldt_b200 import ops
    from ldt_b200.score import _PtrView
    H, dh, Hd, T = 16, 64, 1024, 32
    g = torch.Generator().manual_seed(B)
    M = B * T
    A = torch.randn((M, Hd), generator=g).to(dev).bfloat16()
    Wf = (torch.randn((3 * Hd, Hd), generator=g) / Hd ** 0.5)
    bias = (torch.randn((3 * Hd,), generator=g) * 0.5).to(dev)
    W = Wf.to(dev).bfloat16().contiguous()
    perm = torch.stack([torch.arange(H).view(H, 1) * dh + torch.arange(dh).view(1, dh) + off for off in (0, Hd, 2 * Hd)],
                       dim=1).reshape(-1).to(dev)
    Wp, bp = W[perm].contiguous(), bias[perm].contiguous()
    qkv = torch.empty((M, 3 * Hd), dtype=torch.bfloat16, device=dev)
    o_ref = torch.empty((M, Hd), dtype=torch.bfloat16, device=dev)
    ops.gemm(A, W, bias, qkv, 1)
    ops.attention_nk32(B, H, T, dh, qkv, 3 * Hd, _PtrView(qkv.data_ptr() + 2 * Hd), _PtrView(qkv.data_ptr() + 4 * Hd), 3 * Hd, o_ref)
    o = torch.full((M, Hd), 7.0, dtype=torch.bfloat16, device=dev)
    ops.qkv_attention(B, H, A, Wp, bp, o)
    assert rms_rel_err(o.float(), o_ref.float()) < 2e-3, rms_rel_err(o.float(), o_ref.float())
    # the same fused kernel with the warp-level mma.sync attention arithmetic (the pre-tcgen05 epilogue, kept as a cross-check)
    o_sync = torch.full((M, Hd), 7.0, dtype=torch.bfloat16, device=dev)
    try:
        ops.set_attention_backend(1)
        ops.qkv_attention(B, H, A, Wp, bp, o_sync)
        torch.cuda.synchronize()
    finally:
        ops.set_attention_backend(0)
    assert rms_rel_err(o.float(), o_sync.float()) < 1.5e-3, rms_rel_err(o.float(), o_sync.float())
    assert rel_rms_err(o.float(), o_sync.float()) < 2e-2, rel_rms_err(o.float(), o_sync.float())
    again = torch.empty_like(o)
    ops.qkv_attention(B, H, A, Wp, bp, again)
    assert torch.equal(again, o)     # run-to-run deterministic
    # fp32 formula on the bf16-rounded q/k/v
    q3 = qkv.float()
    qh = q3[:, :Hd].reshape(B, T, H, dh).permute(0, 2, 1, 3)
    kh = q3[:, Hd:2 * Hd].reshape(B, T, H, dh).permute(0, 2, 1, 3)
    vh = q3[:, 2 * Hd:].reshape(B, T, H, dh).permute(0, 2, 1, 3)
    w = ((qh @ kh.transpose(-2, -1)) * dh ** -0.5).softmax(-1)
    ref = (w @ vh).reshape(B, T, Hd)   # [B,H,T,dh] buffer re-read token-major (layout quirk)
    assert rms_rel_err(o.float().view(B, T, Hd), ref) < 4e-3
    assert rel_rms_err(o.float().view(B, T, Hd), ref) < 3e-2


# ------------------------------------------------------------------------------------------------
# SDE update kernel
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("pred", ["ancestral", "reversediffusion", "eulermaruyama", "ddim"])
def test_sde_step_bit_exact_vs_oracle(dev, pred):
    """Teacher-forced: same x, params, z and step scalars in => bit-identical x_next and x_mean as the reference's
    op sequence (oracle restatement of diffusion_continuous.py:141-191 run on the same device with torch ops)."""
    from ldt_b200 import DiffusionVPSDE, ops
    from ldt_b200.sde import _PRED_CODES
    cfg = ns(airplane_config()).sde
    sde = DiffusionVPSDE(cfg, device=dev)
    N = 1000
    coef, ts = sde.step_coefficients(pred, N, 1e-6, False, dev)
    g = torch.Generator().manual_seed(9)
    x = torch.randn((4, 32, 120), generator=g).to(dev)
    prm = torch.randn((4, 32, 120), generator=g).to(dev)
    z = torch.randn((4, 32, 120), generator=g).to(dev)
    osde = O.VPSDE(cfg.beta_start, cfg.beta_end, cfg.sigma2_0, cfg.sample_N)
    # the reference builds its tables on the device (diffusion_continuous.py:649-653): cumprod must run there too
    osde.betas = osde.betas.to(dev)
    osde.alpha = 1.0 - osde.betas
    osde.alphas_cump = osde.alpha.cumprod(dim=0)
    for i in (0, 1, 10, 500, 998, 999):
        step = torch.tensor([i], dtype=torch.int32, device=dev)
        xn, xm = torch.empty_like(x), torch.empty_like(x)
        ops.sde_step(_PRED_CODES[pred], x, prm, z, coef, step, 0, 0, 0, 0, xn, xm)
        t = torch.ones(4, device=dev) * ts[i]
        if pred == "ancestral":
            rn, rm = O.ancestral_step(osde, x, t, prm, z, N)
        elif pred == "reversediffusion":
            rn, rm = O.reverse_diffusion_step(osde, x, t, prm, z, N, 1e-6)
        elif pred == "eulermaruyama":
            rn, rm = O.euler_maruyama_step(osde, x, t, prm, z, N)
        else:
            rn, rm = O.ddim_step(osde, x, t, prm, N)
        assert torch.equal(xm, rm), f"{pred} step {i}: x_mean differs by {(xm - rm).abs().max()}"
        assert torch.equal(xn, rn), f"{pred} step {i}: x differs by {(xn - rn).abs().max()}"


@pytest.mark.parametrize("numel_shape", [(2, 32, 120), (16, 32, 120), (256, 32, 120), (333, 32, 120)])
def test_sde_step_philox_noise_equals_torch_randn_like(dev, numel_shape):
    """With z == NULL the kernel's in-kernel Philox normals must equal torch.randn_like from the same generator
    state -- the reference draws its per-step noise that way (diffusion_continuous.py:160)."""
    from ldt_b200 import ops
    from ldt_b200.sde import torch_randn_launch_geometry
    x = torch.zeros(numel_shape, device=dev)
    prm = torch.zeros_like(x)
    coef = torch.zeros((1, 8), device=dev)
    coef[0, 0], coef[0, 1], coef[0, 2], coef[0, 3] = 1.0, 0.0, 1.0, 1.0  # ancestral with beta -> x_next = z exactly
    gen = torch.cuda.default_generators[0]
    torch.cuda.manual_seed(1234)
    seed, off = gen.initial_seed(), gen.get_offset()
    want = torch.randn_like(x)
    grid, per_call = torch_randn_launch_geometry(x.numel(), dev)
    assert gen.get_offset() - off == per_call, (gen.get_offset() - off, per_call)
    xn = torch.empty_like(x)
    ops.sde_step(0, x, prm, None, coef, None, seed, off, per_call, grid, xn, None)
    assert torch.equal(xn, want), f"max diff {(xn - want).abs().max()}"
    # second draw: offset advanced by per_call, selected through the device-side step index
    want2 = torch.randn_like(x)
    step = torch.ones(1, dtype=torch.int32, device=dev)
    coef2 = coef.repeat(2, 1)
    ops.sde_step(0, x, prm, None, coef2, step, seed, off, per_call, grid, xn, None)
    assert torch.equal(xn, want2)


# ------------------------------------------------------------------------------------------------
# point-set prologue: furthest point sampling, k-NN grouping, per-step conditioning vector
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("b,n,m,rule", [(3, 600, 32, 1e-3), (2, 2048, 2048, 1e-3), (2, 3500, 2048, 1e-3), (1, 8192, 512, -1.0),
                                         (4, 100, 100, -1.0), (2, 1, 1, -1.0), (2, 513, 7, 1e-3)])
def test_fps_bit_exact_vs_oracle(dev, b, n, m, rule):
    from tests.helpers import oracle_fps
    g = torch.Generator().manual_seed(n + m)
    x = torch.randn((b, n, 3), generator=g)
    x = x / x.norm(dim=-1).max(dim=1)[0][:, None, None]      # unit-sphere normalised like ShapeNet_55.py:50-54
    if n >= 100:
        x[:, 10:14] = x[:, 50:54]                              # duplicate points: exact ties
    got = ops.furthest_point_sample(x.to(dev), m, rule)
    assert got.dtype == torch.int32 and got.shape == (b, m)
    assert torch.equal(got.cpu().long(), oracle_fps(x, m, rule))


@pytest.mark.parametrize("b,n,m", [(4, 2048, 32), (3, 2048, 256), (2, 1000, 100), (2, 4000, 64), (5, 257, 257)])
def test_fps_bit_exact_vs_reference_in_tree_cuda_kernel(dev, b, n, m):
    """The reference's OWN furthest-point-sampling kernel (model/functional/src/sampling/sampling.cu:86-167, compiled from the
    source where it lies into oracle/_ref/libref_fps.so) run on the same GPU: identical index sequences.  That kernel has
    no small-norm skip rule, which is ldt_furthest_point_sample with min_sq_norm < 0; it reads channels-first [b,3,n]
    coordinates and a distance scratch initialised to 1e38 (sampling.cpp:52-53).  n = 4000 crosses its 3072-point shared
    buffer.  Exact distance ties (duplicate points) are the one documented difference and do not occur in these clouds.
    (compute-sanitizer racecheck flags a write-after-read race on dists_i[0] inside the reference's kernel; it does not show at
    these sizes -- the comparison has been bit-identical on every run -- but a mismatch here should first be re-run.)"""
    path = os.path.join(ROOT, "oracle", "_ref", "libref_fps.so")
    assert os.path.exists(path), "oracle/_ref/libref_fps.so missing: run `make -C oracle` in the build container"
    R = C.CDLL(path)
    fn = getattr(R, "_Z23furthest_point_samplingiiiPKfPfPi")   # furthest_point_sampling(b, n, m, coords, distances, indices)
    fn.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    fn.restype = None
    g = torch.Generator().manual_seed(b * 1000 + n + m)
    xyz = torch.randn((b, n, 3), generator=g).to(dev)
    coords = xyz.transpose(1, 2).contiguous()
    dist = torch.full((b, n), 1e38, device=dev)
    want = torch.zeros((b, m), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    fn(b, n, m, coords.data_ptr(), dist.data_ptr(), want.data_ptr())   # legacy default stream
    torch.cuda.synchronize()
    got = ops.furthest_point_sample(xyz, m, -1.0)
    assert torch.equal(got, want), (got != want).nonzero()[:4]
    assert int(want[:, 0].abs().max()) == 0 and len(set(want[0].tolist())) == m


def test_fps_rejects_bad_inputs(dev):
    x = torch.randn((2, 64, 3), device=dev)
    with pytest.raises(RuntimeError):
        ops.furthest_point_sample(x, 65)
    with pytest.raises(RuntimeError):
        ops.furthest_point_sample(x.cpu(), 8)
    with pytest.raises(RuntimeError):
        ops.furthest_point_sample(torch.randn((1, 9000, 3), device=dev), 8)


@pytest.mark.parametrize("b,n,s,k", [(3, 600, 32, 8), (2, 2048, 32, 128), (2, 2048, 2048, 16), (1, 50, 50, 50)])
def test_knn_vs_oracle_and_reference_formula(dev, b, n, s, k):
    from tests.helpers import oracle_knn
    g = torch.Generator().manual_seed(n + s + k)
    x = torch.rand((b, n, 3), generator=g)
    c = x[:, torch.randperm(n, generator=g)[:s]].contiguous()
    got = ops.knn_indices(k, x.to(dev), c.to(dev)).cpu().long()
    assert torch.equal(got, oracle_knn(k, x, c))               # order (distance, index) and ties, bit for bit
    # the reference's knn_point (expanded |a|^2+|b|^2-2ab, topk unsorted) selects the same SET wherever the k-th and
    # (k+1)-th distances are separated by more than its fp32 cancellation error
    ref = O.knn_point(k, x, c)
    d = ((c[:, :, None, :] - x[:, None, :, :]) ** 2).sum(-1)
    srt = d.sort(dim=-1)[0]
    clear = (srt[..., k] - srt[..., k - 1] > 1e-5) if k < n else torch.ones(d.shape[:2], dtype=torch.bool)
    same = (got.sort(-1)[0] == ref.sort(-1)[0]).all(-1)
    assert bool((same | ~clear).all()) and float(clear.float().mean()) > 0.9


def test_cond_silu_matches_time_embedding_path(dev):
    """table row + per-sample vector -> SiLU -> bf16 must equal what ldt_time_embedding produces with `extra`."""
    cfg = small_score_cfg()
    from ldt_b200 import Score
    m = Score(cfg)
    m.load_state_dict(O.synth_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, 11))
    m = m.to(dev).eval()
    P = m.packed()
    from ldt_b200.sampler import time_embedding_table
    ts = torch.linspace(1.0, 1e-6, 7, device=dev)
    table = time_embedding_table(m, P, ts)
    B = 5
    extra = torch.randn((B, cfg.t_dim), device=dev)
    step = torch.tensor([4], dtype=torch.int32, device=dev)
    c = torch.empty((B, cfg.t_dim), device=dev)
    sc = torch.empty((B, cfg.t_dim), dtype=torch.bfloat16, device=dev)
    ops.cond_silu(table, step, extra, c, sc)
    ws = m._workspace(B, B, dev)
    m.modulation(P, ws, ts[4].expand(B).contiguous(), extra)
    assert torch.equal(c, ws.c) and torch.equal(sc, ws.sc)


def test_completion_metrics_chamfer_eval_and_fscore(dev):
    """L2_ChamferEval_1000 / F1Score (completion_trainer/Latent_SDE_Trainer.py:41-53) against the float64 brute-force
    NN distances of the oracle, including a pair with no point under the threshold (0/0 -> 0)."""
    from ldt_b200 import metrics
    g = torch.Generator().manual_seed(8)
    a = torch.rand((5, 300, 3), generator=g) * 0.2
    b = a[:, torch.randperm(300, generator=g)[:256]] + 0.01 * torch.randn((5, 256, 3), generator=g)
    b[4] += 5.0                                                        # far away: precision 0 both ways
    d1, _, d2, _ = O.nn_distance_f64(a, b)
    want_cd = (d1.mean() + d2.mean()) * 1000
    got_cd = metrics.L2_ChamferEval_1000(a.to(dev), b.to(dev))
    assert abs(float(got_cd) - float(want_cd)) <= 1e-5 * float(want_cd)
    f, p1, p2 = metrics.F1Score(a.to(dev), b.to(dev))
    wp1, wp2 = (d1.float() < 0.001).float().mean(1), (d2.float() < 0.001).float().mean(1)
    wf = 2 * wp1 * wp2 / (wp1 + wp2)
    wf[torch.isnan(wf)] = 0
    assert torch.allclose(p1.cpu(), wp1, atol=1 / 300 + 1e-6) and torch.allclose(p2.cpu(), wp2, atol=1 / 256 + 1e-6)
    assert torch.allclose(f.cpu(), wf, atol=2e-2) and float(f[4]) == 0.0
