"""Parity of each C-ABI kernel on a B200 (``-m gpu``).  Integer/byte results are bit-exact against
the oracle; floating-point kernels are compared with a plain torch fp32 evaluation of the same op
on the same (bf16-rounded) operands, with the tolerance written at each assert."""
import math

import pytest
import torch

from oracle import samplernn_oracle as O
from tests.helpers import Golden, rel_l2

pytestmark = pytest.mark.gpu

BF16, F32 = torch.bfloat16, torch.float32


@pytest.fixture(scope='module')
def ops():
    if not torch.cuda.is_available():
        pytest.skip('needs a GPU')
    from samplernn_pase_b200 import ops as _ops
    return _ops


def dev(t):
    return t.to('cuda')


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator().manual_seed(seed + sum(shape))
    return (torch.randn(*shape, generator=g) * scale).cuda()


# ------------------------------------------------------------------------------------------------
# quantiser: bit-exact
# ------------------------------------------------------------------------------------------------
def test_quantize_ulaw_bit_exact_vs_reference_chain_on_cuda(ops):
    g = torch.Generator().manual_seed(11)
    x = torch.cat([torch.linspace(-0.99, 0.99, 2_000_003), (torch.rand(6_000_000, generator=g) * 2 - 1) * 0.99,
                   torch.tensor([0.0, -0.0, 1e-8, -1e-8, 0.5, -0.5])]).cuda()
    want = O.quantize_ulaw(x)                       # the reference op chain, executed by torch on CUDA
    got64, got8 = ops.quantize_ulaw(x, want_i64=True, want_u8=True)
    assert torch.equal(got64, want)
    assert torch.equal(got8.long(), want)


def test_quantize_ulaw_vs_cpu_golden(ops):
    """CPU golden vectors (reference on CPU).  torch-CPU divides where torch-CUDA multiplies by the
    reciprocal and uses a different logf, so a handful of boundary values may differ by one level
    (58 of 5e7); the bound here is <= 1 level and < 1e-4 of the samples."""
    g = Golden('quantizer')
    got = ops.quantize_ulaw(g.t('x').cuda())[0].cpu()
    diff = (got - g.t('idx')).abs()
    assert int(diff.max()) <= 1
    assert float((diff > 0).float().mean()) < 1e-4


def test_quantize_ulaw_ragged_and_overflow(ops):
    for n in (1, 2, 3, 5, 1023):
        x = (torch.rand(n) * 2 - 1).mul(0.99).cuda()
        assert torch.equal(ops.quantize_ulaw(x)[0], O.quantize_ulaw(x))
    over = torch.zeros(1, dtype=torch.int32, device='cuda')
    x = torch.tensor([1.0, 0.5, 1.0], device='cuda')
    i64, u8 = ops.quantize_ulaw(x, want_u8=True, overflow=over)
    assert i64.tolist() == [256, O.quantize_ulaw(torch.tensor([0.5]).cuda()).item(), 256]   # SURVEY trap 3
    assert u8.tolist()[0] == 255 and int(over) == 2


def test_quantize_linear_rows(ops):
    x = (torch.rand(7, 4001, generator=torch.Generator().manual_seed(5)) * 2 - 1).cuda()
    assert torch.equal(ops.quantize_linear(x)[0], O.quantize_linear(x))
    g = Golden('quantizer')
    got = ops.quantize_linear(g.t('x_linear').cuda().view(1, -1))[0].view(-1).cpu()
    assert int((got - g.t('idx_linear')).abs().max()) <= 1


def test_dequantize_and_onehot(ops):
    from samplernn_pase_b200.utils import SampleRNNQuantizer
    g = Golden('quantizer')
    q = SampleRNNQuantizer(True, 256)
    idx = torch.arange(257, device='cuda')
    assert torch.equal(q.dequantize(idx).cpu(), g.t('dequant_table'))          # exact, incl. the lost sign
    ql = SampleRNNQuantizer(False, 256)
    assert torch.equal(ql.dequantize(torch.arange(256, device='cuda')).cpu(), g.t('dequant_linear_table'))
    u8 = torch.randint(0, 256, (3, 37), dtype=torch.uint8, device='cuda')
    oh = ops.onehot_rows(u8)
    assert torch.equal(oh.float(), torch.nn.functional.one_hot(u8.long(), 256).float())


# ------------------------------------------------------------------------------------------------
# GEMM
# ------------------------------------------------------------------------------------------------
def _gemm_ref(a, b):
    return a.float() @ b.float().t()


@pytest.mark.parametrize('m,n,k', [(128, 256, 64), (256, 256, 128), (1000, 1024, 3072), (300, 128, 64),
                                   (4100, 3072, 1024), (77, 32, 96), (513, 50, 64)])
def test_gemm_nt_plain(ops, m, n, k):
    a = rnd(m, k).to(BF16)
    b = rnd(n, k, seed=1).to(BF16)
    c = torch.empty(m, n, dtype=F32, device='cuda')
    ops.gemm_nt(a, b, c, m, n, k, k, k, n)
    ref = _gemm_ref(a, b)
    assert rel_l2(c, ref) < 1e-5, rel_l2(c, ref)           # fp32 accumulation of exact bf16 products
    cb = torch.empty(m, round_up8(n), dtype=BF16, device='cuda')
    ops.gemm_nt(a, b, cb, m, n, k, k, k, round_up8(n))
    assert rel_l2(cb[:, :n], ref) < 4e-3                   # bf16 output rounding (2^-9)


def round_up8(n):
    return (n + 7) // 8 * 8


def test_gemm_nt_epilogue_bias_aux_relu_gate(ops):
    m, n, k = 700, 512, 256
    a, b = rnd(m, k).to(BF16), rnd(n, k, seed=1).to(BF16)
    bias = rnd(n, seed=2)
    aux = rnd(m, n, seed=3).to(BF16)
    c = torch.empty(m, n, dtype=F32, device='cuda')
    ops.gemm_nt(a, b, c, m, n, k, k, k, n, bias=bias, aux=aux, ldaux=n, aux_mode=1, relu=True)
    ref = torch.relu(_gemm_ref(a, b) + bias + aux.float())
    assert rel_l2(c, ref) < 1e-5
    ops.gemm_nt(a, b, c, m, n, k, k, k, n, aux=aux, ldaux=n, aux_mode=2)
    ref = _gemm_ref(a, b) * (aux.float() > 0)
    assert rel_l2(c, ref) < 1e-5


@pytest.mark.parametrize('m,n,k', [(256, 1024, 1024), (200, 72, 4096), (512, 4096, 1024), (64, 256, 192), (3, 3072, 1024)])
def test_gemm_nt_small_m_cluster_split_k(ops, m, n, k):
    """M <= 512 takes the 128x64-tile kernel whose K extent is split over a thread-block cluster (generation GEMMs):
    bias + aux (one aux row per `div` output rows) + ReLU, bf16 and fp32 outputs, strided aux / C."""
    div = 4 if m % 4 == 0 else 1
    a, b = rnd(m, k, scale=0.5).to(BF16), rnd(n, k, seed=1, scale=0.5).to(BF16)
    bias = rnd(n, seed=2)
    aux = rnd(m // div, 2 * n, seed=3).to(BF16)
    ref = torch.relu(_gemm_ref(a, b) + bias + aux[:, n:].float().repeat_interleave(div, dim=0))
    c = torch.zeros(m, n + 8, dtype=F32, device='cuda')
    ops.gemm_nt(a, b, c, m, n, k, k, k, n + 8, bias=bias, aux=aux[:, n:], ldaux=2 * n, aux_mode=1, relu=True,
                aux_row_div=div)
    assert rel_l2(c[:, :n], ref) < 1e-5 and float(c[:, n:].abs().max()) == 0.0
    cb = torch.zeros(m, n + 8, dtype=BF16, device='cuda')
    ops.gemm_nt(a, b, cb, m, n, k, k, k, n + 8, bias=bias, aux=aux[:, n:], ldaux=2 * n, aux_mode=1, relu=True,
                aux_row_div=div)
    assert rel_l2(cb[:, :n], ref) < 4e-3 and float(cb[:, n:].float().abs().max()) == 0.0
    c2 = torch.empty(m, n, dtype=F32, device='cuda')
    ops.gemm_nt(a, b, c2, m, n, k, k, k, n)                # plain; twice: the result is deterministic
    c3 = torch.empty(m, n, dtype=F32, device='cuda')
    ops.gemm_nt(a, b, c3, m, n, k, k, k, n)
    assert rel_l2(c2, _gemm_ref(a, b)) < 1e-5 and torch.equal(c2, c3)


@pytest.mark.parametrize('m,n,k,batch', [(5000, 1024, 256, 1), (777, 200, 64, 3), (148 * 128 + 5, 512, 128, 1)])
def test_gemm_nt_fused_column_sums(ops, m, n, k, batch):
    """colsum: the bias gradient (column sums over all rows and batches of the fp32 epilogue result) comes out of the
    same launch; rows beyond m and columns beyond n contribute nothing."""
    a = rnd(batch, m, k).to(BF16)
    b = rnd(n, k, seed=1).to(BF16)
    aux = rnd(batch, m, n, seed=3).to(BF16)
    c = torch.empty(batch, m, n, dtype=BF16, device='cuda')
    cs = torch.zeros(n, dtype=F32, device='cuda')
    ops.gemm_nt(a, b, c, m, n, k, k, k, n, batch=batch, a_bs=m * k, c_bs=m * n, aux=aux, ldaux=n, aux_bs=m * n, aux_mode=2,
                colsum=cs)
    ref = (a.float() @ b.float().t()) * (aux.float() > 0)
    assert rel_l2(c, ref) < 4e-3
    assert rel_l2(cs, ref.sum(dim=(0, 1))) < 1e-4, rel_l2(cs, ref.sum(dim=(0, 1)))
    cs2 = torch.zeros(n, dtype=F32, device='cuda')                 # plain epilogue, fp32 output, bias
    bias = rnd(n, seed=5)
    c32 = torch.empty(batch, m, n, dtype=F32, device='cuda')
    ops.gemm_nt(a, b, c32, m, n, k, k, k, n, batch=batch, a_bs=m * k, c_bs=m * n, bias=bias, colsum=cs2)
    ref2 = a.float() @ b.float().t() + bias
    assert rel_l2(cs2, ref2.sum(dim=(0, 1))) < 1e-4


@pytest.mark.parametrize('m,n,k1,k2,batch', [(1000, 512, 128, 200, 1), (700, 256, 1024, 64, 2), (130, 64, 64, 16, 1)])
def test_gemm_nt_two_a_operands_concatenated_along_k(ops, m, n, k1, k2, batch):
    """A = [A1 | A2] without materialising the concatenation (comb_layer's [one-hot windows | upper] input)."""
    a1 = rnd(batch, m, k1, scale=0.5).to(BF16)
    a2 = rnd(batch, m, round_up8(k2), scale=0.5, seed=1).to(BF16)
    b = rnd(n, k1 + k2, seed=2, scale=0.5).to(BF16)
    bp = torch.zeros(n, round_up8(k1 + k2), dtype=BF16, device='cuda')
    bp[:, :k1 + k2] = b
    c = torch.empty(batch, m, n, dtype=F32, device='cuda')
    ops.gemm_nt(a1, bp, c, m, n, k1 + k2, k1, bp.shape[1], n, batch=batch, a_bs=m * k1, c_bs=m * n,
                a2=a2, lda2=a2.shape[2], a2_bs=m * a2.shape[2], k1=k1)
    ref = torch.cat([a1.float(), a2[..., :k2].float()], dim=2) @ b.float().t()
    assert rel_l2(c, ref) < 1e-5, rel_l2(c, ref)


def test_gemm_nt_batched_overlapping_rows_and_strided_c(ops):
    """The sample-level contraction: A rows are overlapping windows of a (B, W, Q) one-hot buffer."""
    bsz, rf, r0, q, h = 3, 200, 4, 256, 64
    w = rf + r0 - 1
    idx = torch.randint(0, q, (bsz, w), dtype=torch.uint8, device='cuda')
    onehot = ops.onehot_rows(idx)
    table = rnd(h, r0 * q, seed=4).to(BF16)
    cat = torch.zeros(bsz * rf, 3 * h, dtype=BF16, device='cuda')
    ops.gemm_nt(onehot, table, cat, rf, h, r0 * q, q, r0 * q, 3 * h, batch=bsz, a_bs=w * q, c_bs=rf * 3 * h)
    win = onehot.float().unfold(1, r0, 1).permute(0, 1, 3, 2).reshape(bsz * rf, r0 * q)
    ref = win @ table.float().t()
    assert rel_l2(cat[:, :h], ref) < 4e-3
    assert float(cat[:, h:].abs().max()) == 0.0


def test_gemm_nt_fold_is_the_upsample_layout(ops):
    bsz, t, h, r = 2, 37, 64, 4
    hall = rnd(bsz, t + 1, h).to(BF16)
    wu = rnd(r * h, h, seed=1).to(BF16)
    bias = rnd(r * h, seed=2)
    up = torch.empty(bsz, t * r, h, dtype=BF16, device='cuda')
    ops.gemm_nt(hall[:, 1:], wu, up, t, r * h, h, h, h, r * h, batch=bsz, a_bs=(t + 1) * h, c_bs=t * r * h, bias=bias)
    ref = (hall[:, 1:].float() @ wu.float().t() + bias).reshape(bsz, t * r, h)
    assert rel_l2(up, ref) < 4e-3
    # n_fold: same values written with a leading dimension of 3H (straight into the concat buffer)
    cat = torch.zeros(bsz * t * r, 3 * h, dtype=BF16, device='cuda')
    ops.gemm_nt(hall[:, 1:], wu, cat[:, 2 * h:], t, r * h, h, h, h, 3 * h, batch=bsz, a_bs=(t + 1) * h,
                c_bs=t * r * 3 * h, bias=bias, n_fold=h)
    assert rel_l2(cat[:, 2 * h:], ref.reshape(-1, h)) < 4e-3


@pytest.mark.parametrize('m,n,k', [(128, 128, 64), (256, 256, 1000), (1024, 3072, 5000), (56, 64, 333), (3072, 1024, 20000)])
def test_gemm_tn(ops, m, n, k):
    a = rnd(k, m).to(BF16)
    b = rnd(k, n, seed=1).to(BF16)
    c = torch.zeros(m, n, dtype=F32, device='cuda')
    ops.gemm_tn(a, b, c, m, n, k, m, n, n)
    ref = a.float().t() @ b.float()
    assert rel_l2(c, ref) < 2e-5, rel_l2(c, ref)
    ops.gemm_tn(a, b, c, m, n, k, m, n, n)                  # accumulates
    assert rel_l2(c, 2 * ref) < 2e-5


def test_gemm_tn_batched_with_row_offset(ops):
    bsz, rf, r0, q, h = 3, 150, 4, 256, 64
    w = rf + r0 - 1
    idx = torch.randint(0, q, (bsz, w), dtype=torch.uint8, device='cuda')
    onehot = ops.onehot_rows(idx)
    de = rnd(bsz, rf, h).to(BF16)
    g = torch.zeros(q, r0 * h, dtype=F32, device='cuda')
    for k in range(r0):
        ops.gemm_tn(onehot, de, g[:, k * h:], q, h, rf, q, h, r0 * h, batch=bsz, a_bs=w * q, b_bs=rf * h, a_off=k)
    for k in range(r0):
        ref = torch.einsum('bjq,bjo->qo', onehot[:, k:k + rf].float(), de.float())
        assert rel_l2(g[:, k * h:(k + 1) * h], ref) < 2e-5


@pytest.mark.parametrize('m,k', [(128, 64), (1000, 1024), (333, 32)])
def test_gemm_nll_all_modes(ops, m, k):
    a = rnd(m, k, scale=0.5).to(BF16)
    w = rnd(256, k, scale=0.2, seed=1).to(BF16)
    bias = rnd(256, seed=2)
    tgt = torch.randint(0, 256, (m,), dtype=torch.uint8, device='cuda')
    logits = _gemm_ref(a, w) + bias
    logp_ref = torch.log_softmax(logits, dim=1)
    lse = torch.empty(m, device='cuda'); lpt = torch.empty(m, device='cuda')
    ops.gemm_nll(0, a, w, bias, tgt, m, k, k, k, lse=lse, logp_target=lpt)
    assert float((lse - torch.logsumexp(logits, 1)).abs().max()) < 1e-4
    assert float((lpt - logp_ref.gather(1, tgt.long()[:, None])[:, 0]).abs().max()) < 1e-4
    logp = torch.empty(m, 256, device='cuda')
    ops.gemm_nll(1, a, w, bias, tgt, m, k, k, k, lse=lse, logp_target=lpt, logp=logp)
    assert float((logp - logp_ref).abs().max()) < 1e-4
    rg = rnd(m, seed=3)
    dl = torch.empty(m, 256, dtype=BF16, device='cuda')
    ops.gemm_nll(2, a, w, bias, tgt, m, k, k, k, row_grad=rg, dlogits=dl)
    ref = rg[:, None] * (torch.nn.functional.one_hot(tgt.long(), 256).float() - logp_ref.exp())
    assert rel_l2(dl, ref) < 4e-3
    dl1 = torch.empty_like(dl)                                   # single pass with the forward's lse: same result
    ops.gemm_nll(2, a, w, bias, tgt, m, k, k, k, row_grad=rg, dlogits=dl1, lse=lse)
    assert rel_l2(dl1, ref) < 4e-3 and rel_l2(dl1, dl) < 1e-3
    g = rnd(m, 256, seed=4)
    ops.gemm_nll(3, a, w, bias, tgt, m, k, k, k, g=g, dlogits=dl)
    ref = g - logp_ref.exp() * g.sum(1, keepdim=True)
    assert rel_l2(dl, ref) < 4e-3
    ops.gemm_nll(3, a, w, bias, tgt, m, k, k, k, g=g, dlogits=dl1, lse=lse)
    assert rel_l2(dl1, ref) < 4e-3 and rel_l2(dl1, dl) < 1e-3


# ------------------------------------------------------------------------------------------------
# persistent GRU
# ------------------------------------------------------------------------------------------------
def _gru_ref(gi, w_hh, b_hh, h0):
    """fp32 evaluation of the same recurrence with the matmul operand rounded to bf16 like the kernel."""
    bsz, t, h3 = gi.shape
    h = h3 // 3
    hs, gates = [], []
    cur = h0.clone()
    for s in range(t):
        gh = cur.to(BF16).float() @ w_hh.float().t() + b_hh
        r = torch.sigmoid(gi[:, s, :h] + gh[:, :h])
        z = torch.sigmoid(gi[:, s, h:2 * h] + gh[:, h:2 * h])
        hn = gh[:, 2 * h:]
        n = torch.tanh(gi[:, s, 2 * h:] + r * hn)
        cur = (1 - z) * n + z * cur
        hs.append(cur)
        gates.append(torch.cat([r, z, n, hn], 1))
    return torch.stack(hs, 1), torch.stack(gates, 1), cur


@pytest.mark.parametrize('bsz,t,h', [(3, 5, 32), (8, 40, 64), (64, 33, 1024), (100, 9, 256), (130, 6, 64)])
def test_gru_forward(ops, bsz, t, h):
    gi = rnd(bsz, t, 3 * h).to(BF16)
    w_hh = rnd(3 * h, h, scale=1 / math.sqrt(h), seed=1).to(BF16)
    b_hh = rnd(3 * h, scale=0.1, seed=2)
    h0 = rnd(bsz, h, scale=0.5, seed=3)
    h_ext = torch.zeros(t + 1, bsz, h, dtype=BF16, device='cuda')      # time-major exchange buffer
    h_ext[0] = h0.to(BF16)
    hall = torch.zeros(bsz, t, h, dtype=BF16, device='cuda')
    h_state = h0.clone()
    gates = torch.empty(bsz * t, 4 * h, dtype=BF16, device='cuda')
    ops.gru_forward(gi.view(bsz * t, 3 * h), w_hh, b_hh, h_ext, hall, h_state, gates, bsz, t, h)
    hs, gref, hT = _gru_ref(gi.float(), w_hh, b_hh, h0)
    assert float((hall.float() - hs).abs().max()) < 2e-2               # bf16 storage of values in (-1,1)
    assert torch.equal(h_ext[1:].transpose(0, 1), hall)                # both copies of h_t agree
    assert float((h_state - hT).abs().max()) < 1e-2
    assert float((gates.view(bsz, t, 4 * h).float() - gref).abs().max()) < 3e-2


@pytest.mark.parametrize('bsz,t,h', [(3, 5, 32), (8, 40, 64), (64, 17, 1024), (100, 9, 256)])
def test_gru_backward(ops, bsz, t, h):
    gi = rnd(bsz, t, 3 * h).to(BF16)
    w_hh = rnd(3 * h, h, scale=1 / math.sqrt(h), seed=1).to(BF16)
    b_hh = rnd(3 * h, scale=0.1, seed=2)
    h0 = rnd(bsz, h, scale=0.5, seed=3)
    dh_out = rnd(bsz, t, h, scale=0.1, seed=4).to(BF16)
    # fp32 autograd reference of the same recurrence
    gi_r = gi.float().requires_grad_(True)
    h0_r = h0.clone().requires_grad_(True)
    cur, outs = h0_r, []
    for s in range(t):
        gh = cur @ w_hh.float().t() + b_hh
        r = torch.sigmoid(gi_r[:, s, :h] + gh[:, :h])
        z = torch.sigmoid(gi_r[:, s, h:2 * h] + gh[:, h:2 * h])
        n = torch.tanh(gi_r[:, s, 2 * h:] + r * gh[:, 2 * h:])
        cur = (1 - z) * n + z * cur
        outs.append(cur)
    (torch.stack(outs, 1) * dh_out.float()).sum().backward()
    # kernel path
    h_ext = torch.zeros(t + 1, bsz, h, dtype=BF16, device='cuda')
    h_ext[0] = h0.to(BF16)
    h_state = h0.clone()
    gates = torch.empty(bsz * t, 4 * h, dtype=BF16, device='cuda')
    ops.gru_forward(gi.view(bsz * t, 3 * h), w_hh, b_hh, h_ext, None, h_state, gates, bsz, t, h)
    dgi = torch.empty(bsz * t, 3 * h, dtype=BF16, device='cuda')
    dgh = torch.empty(bsz * t, 3 * h, dtype=BF16, device='cuda')
    dh0 = torch.empty(bsz, h, dtype=F32, device='cuda')
    db_ih = torch.full((3 * h,), 0.5, dtype=F32, device='cuda')       # accumulated INTO: start from a known value
    db_hh = torch.zeros(3 * h, dtype=F32, device='cuda')
    ops.gru_backward(w_hh.t().contiguous(), h_ext, gates, dh_out.view(bsz * t, h), dgi, dgh, dh0, bsz, t, h, db_ih, db_hh)
    assert rel_l2(dgi.view(bsz, t, 3 * h), gi_r.grad) < 3e-2, rel_l2(dgi.view(bsz, t, 3 * h), gi_r.grad)
    assert rel_l2(dh0, h0_r.grad) < 3e-2, rel_l2(dh0, h0_r.grad)
    # fused bias gradients = column sums of the gate gradients the kernel wrote (fp32 sums of the unrounded values)
    assert rel_l2(db_ih - 0.5, dgi.float().sum(0)) < 1e-2, rel_l2(db_ih - 0.5, dgi.float().sum(0))
    assert rel_l2(db_hh, dgh.float().sum(0)) < 1e-2, rel_l2(db_hh, dgh.float().sum(0))


def test_gru_grid_handshake_stress_bit_exact(ops):
    """The recurrence's data path has no atomics, so repeated launches must be BIT-identical; a stale or incomplete read
    across the grid-wide handshake would show up as a mismatch.  Default protocol: relaxed counter increment, TMA read
    validated through the NaN sentinel (repeated when an element had not arrived); reference: the strict protocol
    (tuning flag 16: release increment + acquire fence).  40 launches x 2000 steps x 128 CTAs = 10 M handshakes on an
    idle GPU, then again while another stream saturates HBM/L2 with copies and runs GEMMs on the SMs the recurrence
    leaves free - the condition under which a data store is most likely to arrive after the counter."""
    bsz, t, h = 64, 2000, 1024
    gi = rnd(bsz, t, 3 * h).to(BF16)
    w_hh = rnd(3 * h, h, scale=1 / math.sqrt(h), seed=1).to(BF16)
    b_hh = rnd(3 * h, scale=0.1, seed=2)
    h0 = rnd(bsz, h, scale=0.5, seed=3)
    dh_out = rnd(bsz, t, h, scale=0.1, seed=4).to(BF16)

    def run(flags):
        ops.gru_tuning_flags = flags
        try:
            h_ext = torch.zeros(t + 1, bsz, h, dtype=BF16, device='cuda')
            h_ext[0] = h0.to(BF16)
            hall = torch.zeros(bsz * t, h, dtype=BF16, device='cuda')
            h_state = h0.clone()
            gates = torch.empty(bsz * t, 4 * h, dtype=BF16, device='cuda')
            ops.gru_forward(gi.view(bsz * t, 3 * h), w_hh, b_hh, h_ext, hall, h_state, gates, bsz, t, h)
            dgi = torch.empty(bsz * t, 3 * h, dtype=BF16, device='cuda')
            dgh = torch.empty(bsz * t, 3 * h, dtype=BF16, device='cuda')
            dh0 = torch.empty(bsz, h, dtype=F32, device='cuda')
            ops.gru_backward(w_hh.t().contiguous(), h_ext, gates, dh_out.view(bsz * t, h), dgi, dgh, dh0, bsz, t, h)
            return hall, h_state, dgi, dh0
        finally:
            ops.gru_tuning_flags = 0

    def retries():
        return int(ops.gru_last_sync[32])

    ref = run(16)                       # flag 16: strict protocol
    seen = 0
    for i in range(20):
        for flags in (0, 16):
            got = run(flags)
            seen += retries() if flags == 0 else 0
            assert all(torch.equal(a, b) for a, b in zip(got, ref)), (i, flags)
    # under load: a side stream streams 1 GB copies and runs tensor-core GEMMs capped to the 20 SMs the 128-CTA
    # recurrence leaves free (the GEMM grid must stay below 148 - 128 CTAs: the recurrence is a cooperative launch)
    side = torch.cuda.Stream()
    big_a = torch.empty(1 << 28, dtype=torch.float32, device='cuda')
    big_b = torch.empty_like(big_a)
    ga = rnd(16384, 1024).to(BF16); gb = rnd(1024, 1024).to(BF16)
    gc = torch.empty(16384, 1024, dtype=BF16, device='cuda')
    seen_loaded = 0
    for i in range(6):
        with torch.cuda.stream(side):
            ops.gemm_max_ctas = 16
            try:
                for _ in range(40):
                    big_b.copy_(big_a, non_blocking=True)
                    ops.gemm_nt(ga, gb, gc, 16384, 1024, 1024, 1024, 1024, 1024)
            finally:
                ops.gemm_max_ctas = 0
        got = run(0)
        seen_loaded += retries()
        side.synchronize()
        assert all(torch.equal(a, b) for a, b in zip(got, ref)), ('loaded', i)
    import os
    try:
        with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out', 'parity_report.txt'), 'a') as f:
            f.write(f'gru handshake stress: bit-exact; rejected+repeated exchange attempts (of 160 K timesteps x 128 CTAs idle, '
                    f'48 K under load): idle {seen}, loaded {seen_loaded}\n')
    except OSError:
        pass


def test_gru_and_lstm_with_16_units_per_cta(ops):
    """The half-grid variant of the recurrent kernels (H/16 CTAs) must give the same results."""
    ops.gru_units_per_cta = 16
    try:
        test_gru_forward(ops, 64, 33, 1024)
        test_gru_backward(ops, 64, 17, 1024)
        test_gru_forward(ops, 8, 40, 128)
        test_gru_backward(ops, 8, 40, 128)
        test_lstm_forward_backward(ops, 64, 17, 1024)
        test_lstm_forward_backward(ops, 8, 20, 128)
    finally:
        ops.gru_units_per_cta = 8


@pytest.mark.parametrize('bsz,t,h', [(3, 5, 32), (8, 20, 64), (64, 17, 1024), (70, 6, 128)])
def test_lstm_forward_backward(ops, bsz, t, h):
    """LSTM extension (BASELINE config 3; no reference counterpart): torch.nn.LSTM semantics, checked
    against an fp32 autograd evaluation of the same recurrence."""
    gi = rnd(bsz, t, 4 * h).to(BF16)
    w_hh = rnd(4 * h, h, scale=1 / math.sqrt(h), seed=1).to(BF16)
    b_hh = rnd(4 * h, scale=0.1, seed=2)
    h0 = rnd(bsz, h, scale=0.5, seed=3)
    c0 = rnd(bsz, h, scale=0.5, seed=5)
    dh_out = rnd(bsz, t, h, scale=0.1, seed=4).to(BF16)
    gi_r = gi.float().requires_grad_(True)
    h0_r, c0_r = h0.clone().requires_grad_(True), c0.clone().requires_grad_(True)
    hh, cc, outs = h0_r, c0_r, []
    for s in range(t):
        g = gi_r[:, s] + hh @ w_hh.float().t() + b_hh
        i_, f_, g_, o_ = torch.sigmoid(g[:, :h]), torch.sigmoid(g[:, h:2 * h]), torch.tanh(g[:, 2 * h:3 * h]), torch.sigmoid(g[:, 3 * h:])
        cc = f_ * cc + i_ * g_
        hh = o_ * torch.tanh(cc)
        outs.append(hh)
    ref_out = torch.stack(outs, 1)
    (ref_out * dh_out.float()).sum().backward()
    h_ext = torch.zeros(t + 1, bsz, h, dtype=BF16, device='cuda')
    h_ext[0] = h0.to(BF16)
    hall = torch.zeros(bsz, t, h, dtype=BF16, device='cuda')
    h_state, c_state = h0.clone(), c0.clone()
    gates = torch.empty(bsz * t, 5 * h, dtype=BF16, device='cuda')
    ops.lstm_forward(gi.view(bsz * t, 4 * h), w_hh, b_hh, h_ext, hall, h_state, c_state, gates, bsz, t, h)
    assert float((hall.float() - ref_out.detach()).abs().max()) < 3e-2
    assert float((h_state - hh.detach()).abs().max()) < 2e-2 and float((c_state - cc.detach()).abs().max()) < 3e-2
    dgi = torch.empty(bsz * t, 4 * h, dtype=BF16, device='cuda')
    dgh = torch.empty(t * bsz, 4 * h, dtype=BF16, device='cuda')
    dh0 = torch.empty(bsz, h, dtype=F32, device='cuda')
    dc0 = torch.empty(bsz, h, dtype=F32, device='cuda')
    ops.lstm_backward(w_hh.t().contiguous(), h_ext, gates, c0, dh_out.view(bsz * t, h), dgi, dgh, dh0, dc0, bsz, t, h)
    assert rel_l2(dgi.view(bsz, t, 4 * h), gi_r.grad) < 3e-2, rel_l2(dgi.view(bsz, t, 4 * h), gi_r.grad)
    assert rel_l2(dh0, h0_r.grad) < 3e-2 and rel_l2(dc0, c0_r.grad) < 3e-2
    assert torch.equal(dgh.view(t, bsz, 4 * h).transpose(0, 1).contiguous().view(bsz * t, 4 * h), dgi)


# ------------------------------------------------------------------------------------------------
# small kernels
# ------------------------------------------------------------------------------------------------
def test_weight_prep_and_backward(ops):
    r, a, b = 48, 40, 4
    v = rnd(r, a, b).requires_grad_(True)
    g = (rnd(r, seed=1).abs() + 0.5).requires_grad_(True)
    w_ref = O.weight_norm(g.view(r, 1, 1), v)
    out1 = torch.empty(b * a, r, dtype=BF16, device='cuda')          # upsample layout [(j*A + o), i]
    out2 = torch.empty(r, b * a, dtype=BF16, device='cuda')
    inv = torch.empty(r, device='cuda')
    ops.weight_prep(v.detach(), g.detach(), (r, a, b), out1, (1, r, a * r), out2, (b * a, 1, a), inv)
    want = w_ref.detach().permute(2, 1, 0).reshape(b * a, r)
    assert rel_l2(out1, want) < 4e-3 and rel_l2(out2, want.t()) < 4e-3
    dw = rnd(b * a, r, seed=2)
    (w_ref.permute(2, 1, 0).reshape(b * a, r) * dw).sum().backward()
    dv, dg = ops.weight_prep_bwd(dw, (1, r, a * r), v.detach(), g.detach(), inv, (r, a, b))
    assert rel_l2(dv, v.grad) < 1e-5 and rel_l2(dg, g.grad) < 1e-5
    plain = torch.empty(r, a * b, dtype=BF16, device='cuda')
    ops.weight_prep(v.detach(), None, (r, a * b, 1), plain, (a * b, 1, 0))
    assert torch.equal(plain, v.detach().reshape(r, a * b).to(BF16))


def test_colsum_repeat_cast(ops):
    x = rnd(1000, 200).to(BF16)
    assert rel_l2(ops.colsum(x, 1000, 200, 200), x.float().sum(0)) < 1e-5
    src = rnd(30, 64).to(BF16)
    out = torch.zeros(30 * 5, 192, dtype=BF16, device='cuda')
    ops.repeat_rows(src, 30, 64, 64, 5, out[:, 64:], 192)
    assert torch.equal(out[:, 64:128], src.repeat_interleave(5, 0)) and float(out[:, :64].abs().max()) == 0
    back = torch.empty(30, 64, dtype=BF16, device='cuda')
    ops.repeat_rows_bwd(out[:, 64:], 30, 64, 192, 5, back, 64)
    assert rel_l2(back, 5 * src.float()) < 4e-3
    f = rnd(17, 50)
    padded = torch.empty(17, 56, dtype=BF16, device='cuda')
    ops.pad_cast_bf16(f, 17, 50, 50, padded, 56, 56)
    assert torch.equal(padded[:, :50], f.to(BF16)) and float(padded[:, 50:].abs().max()) == 0


def test_tier_and_mixer_inputs(ops):
    from samplernn_pase_b200.utils import SampleRNNQuantizer
    q = SampleRNNQuantizer(True, 256)
    bsz, l, c, fs, fs_top, rf = 3, 5, 50, 4, 16, 80
    t = rf // fs
    xq = torch.randint(0, 256, (bsz, rf + fs_top - 1), dtype=torch.uint8, device='cuda')
    conds = rnd(bsz, l, c)
    lut = q.lut('cuda')
    kp = 56
    out = ops.tier_input(xq, fs_top - fs, lut, None, conds, bsz, t, fs, kp)
    frames = lut[xq[:, fs_top - fs: fs_top - fs + rf].long()].view(bsz, t, fs)
    want = torch.cat([frames, conds.repeat_interleave(t // l, 1), torch.zeros(bsz, t, kp - fs - c, device='cuda')], 2)
    assert torch.equal(out.view(bsz, t, kp), want.to(BF16))
    out2 = ops.tier_input(None, 0, None, frames.contiguous(), conds, bsz, t, fs, kp)
    assert torch.equal(out2, out)
    dc = torch.zeros(bsz, l, c, device='cuda')
    ops.tier_input_bwd(out, bsz, t, fs, l, c, kp, dc)
    assert rel_l2(dc, out.view(bsz, l, t // l, kp)[..., fs:fs + c].float().sum(2)) < 1e-5
    table = rnd(7, 15, seed=1)
    spk = torch.tensor([3, 0, 6], dtype=torch.int32, device='cuda')
    utt = rnd(bsz, l, 43, seed=2)
    mix = ops.mixer_input(utt, table, spk, 64)
    want = torch.cat([table[spk.long()][:, None].expand(bsz, l, 15), utt, torch.zeros(bsz, l, 6, device='cuda')], 2)
    assert torch.equal(mix.view(bsz, l, 64), want.to(BF16))


def test_state_select_and_masked_mean(ops):
    bsz, h = 5, 64
    carried, h0 = rnd(bsz, h), rnd(h, seed=1)
    use = torch.tensor([1, 0, 1, 0, 0], dtype=torch.uint8, device='cuda')
    hs = ops.state_select(carried, h0, use, bsz, h)
    assert torch.equal(hs, torch.where(use.bool()[:, None], carried, h0[None].expand(bsz, h)))
    dh = rnd(bsz, h, seed=2)
    assert rel_l2(ops.state_select_bwd(dh, use, bsz, h), dh[~use.bool()].sum(0)) < 1e-6
    lp = -rnd(5 * 40, seed=3).abs()
    valid = torch.tensor([1, 1, 0, 1, 0], dtype=torch.uint8, device='cuda')
    out = ops.masked_nll_mean(lp, valid, 40)
    keep = valid.bool().repeat_interleave(40)
    assert abs(float(out[0]) + float(lp[keep].mean())) < 1e-5 and int(out[1]) == 120


def test_adam_clipped_matches_reference_golden(ops):
    g = Golden('adam_clipped')
    for key in 'ab':
        w = g.t(f'w0_{key}').cuda().contiguous()
        m, v = torch.zeros_like(w), torch.zeros_like(w)
        for s in range(3):
            ops.adam_clipped(w, g.t(f'g{s}_{key}').cuda().contiguous(), m, v, 1e-3, 0.9, 0.999, 1e-8, s + 1)
        assert float((w.cpu() - g.t(f'w3_{key}')).abs().max()) < 1e-6


# ------------------------------------------------------------------------------------------------
# generation kernels (model.py:289-351)
# ------------------------------------------------------------------------------------------------
def test_embed_sum_matches_gather_and_sum(ops):
    """srnn_embed_sum == relu(sum of the r0 selected table rows + pre), fp32 accumulate, one bf16 rounding."""
    b, r0, q, h, fs = 37, 4, 256, 192, 16
    table = rnd(r0 * q, h, scale=0.5).to(BF16)
    win = torch.randint(0, 256, (b, fs), generator=torch.Generator().manual_seed(2)).to(torch.uint8).cuda()
    pre = rnd(b, 3, h, scale=0.5, seed=1).to(BF16)
    out = torch.empty(b, h, dtype=BF16, device='cuda')
    ops.embed_sum(table, win[:, fs - r0:], fs, b, r0, q, h, pre[:, 1], 3 * h, True, out, h)
    idx = win[:, fs - r0:].long() + torch.arange(r0, device='cuda') * q
    ref = pre[:, 1].float()
    for k in range(r0):
        ref = ref + table[idx[:, k]].float()
    ref = torch.relu(ref).to(BF16)
    assert torch.equal(out, ref)
    ops.embed_sum(table, win[:, fs - r0:], fs, b, r0, q, h, None, 0, False, out, h)      # no pre, no relu
    ref = sum(table[idx[:, k]].float() for k in range(r0)).to(BF16)
    assert torch.equal(out, ref)


def test_sample_categorical_inverse_cdf_argmax_and_window(ops):
    b, q, fs = 300, 256, 16
    g = torch.Generator().manual_seed(7)
    logits = torch.randn(b, q, generator=g) * 3
    logits[5, 40:] = -float('inf')                                            # mass on a prefix only
    logits[6, :200] = -float('inf')                                           # mass on a suffix only
    logp = torch.log_softmax(logits, dim=1).cuda()
    u = torch.rand(b, generator=g)
    u[0], u[1] = 0.0, 0.99999994                                              # the ends of [0, 1)
    win = torch.randint(0, 256, (b, fs), generator=g).to(torch.uint8).cuda()
    win0 = win.clone()
    out = torch.zeros(b, 3, dtype=torch.uint8, device='cuda')
    ops.sample_categorical(logp, b, q, u.cuda(), win, fs, out[:, 1], 3)
    pick = out[:, 1].long().cpu()
    assert torch.equal(win[:, :-1], win0[:, 1:]) and torch.equal(win[:, -1], out[:, 1])    # shifted window
    assert bool((out[:, 0] == 0).all()) and bool((out[:, 2] == 0).all())                    # strided store
    # the pick is the inverse CDF of u up to fp32 summation error at the class boundaries
    cdf = torch.exp(logp.double().cpu()).cumsum(1)
    tgt = (u.double() * cdf[:, -1])
    lo = torch.where(pick > 0, cdf.gather(1, (pick - 1).clamp(min=0)[:, None])[:, 0], torch.zeros(b, dtype=torch.float64))
    hi = cdf.gather(1, pick[:, None])[:, 0]
    assert bool(((tgt >= lo - 1e-5) & (tgt <= hi + 1e-5)).all())
    assert bool((torch.exp(logp.cpu()).gather(1, pick[:, None]) > 0).all())                # never a zero-mass class
    assert pick[5] < 40 and pick[6] >= 200
    # arg-max mode, lowest index on ties
    logp2 = logp.clone()
    logp2[3, 17] = logp2[3, 99] = 1.0
    ops.sample_categorical(logp2, b, q, None, None, 0, out[:, 0], 3)
    assert torch.equal(out[:, 0].long().cpu(), logp2.cpu().argmax(dim=1)) and int(out[3, 0]) == 17
    # raw logits in: the log-softmax is taken by the kernel (and returned), the same uniforms give the same picks
    lp = torch.zeros(b, 2, q, dtype=F32, device='cuda')
    raw = (logits + 3.0).cuda()
    ops.sample_categorical(raw, b, q, u.cuda(), None, 0, out[:, 2], 3, normalise=True, logp_out=lp[:, 1])
    finite = torch.isfinite(logp)
    assert float((lp[:, 1][finite] - logp[finite]).abs().max()) <= 2e-6 and bool((lp[:, 1][~finite] == -float('inf')).all())
    assert float((out[:, 2] != out[:, 1]).float().mean()) <= 0.01 and float(lp[:, 0].abs().max()) == 0.0
    # empirical distribution of many draws from one row
    n = 200_000
    row = torch.log_softmax(torch.randn(1, q, generator=g) * 2, dim=1)
    draws = torch.empty(n, dtype=torch.uint8, device='cuda')
    ops.sample_categorical(row.expand(n, q).contiguous().cuda(), n, q, torch.rand(n, generator=g).cuda(), None, 0, draws, 1)
    freq = torch.bincount(draws.long().cpu(), minlength=q).double() / n
    p = torch.exp(row[0].double())
    assert float((freq - p).abs().max()) <= 5 * float(torch.sqrt(p.max() / n)) + 1e-4


def test_sample_categorical_device_side_philox_draws(ops):
    """Device-side draws (rng_state = {seed, step, 0}): the empirical distribution matches, the same seed reproduces the
    same picks, every launch advances the step (so replays of a captured graph draw fresh numbers), different
    utterances at one step and one utterance at different steps draw different uniforms."""
    q, n = 256, 200_000
    g = torch.Generator().manual_seed(11)
    row = torch.log_softmax(torch.randn(1, q, generator=g) * 2, dim=1)
    rows = row.expand(n, q).contiguous().cuda()
    state = torch.tensor([1234567, 0, 0], dtype=torch.int64, device='cuda')
    d1 = torch.empty(n, dtype=torch.uint8, device='cuda')
    ops.sample_categorical(rows, n, q, None, None, 0, d1, 1, rng_state=state)
    assert state.tolist() == [1234567, 1, 0]                                   # step advanced, done-counter reset
    freq = torch.bincount(d1.long().cpu(), minlength=q).double() / n
    p = torch.exp(row[0].double())
    assert float((freq - p).abs().max()) <= 5 * float(torch.sqrt(p.max() / n)) + 1e-4
    d2 = torch.empty_like(d1)
    ops.sample_categorical(rows, n, q, None, None, 0, d2, 1, rng_state=state)  # next step: different draws
    assert state.tolist() == [1234567, 2, 0]
    assert 0.5 < float((d1 != d2).float().mean()) < 1.0
    state2 = torch.tensor([1234567, 0, 0], dtype=torch.int64, device='cuda')
    d3 = torch.empty_like(d1)
    ops.sample_categorical(rows, n, q, None, None, 0, d3, 1, rng_state=state2)  # same seed and step: same picks
    assert torch.equal(d1, d3)
    state3 = torch.tensor([7654321, 0, 0], dtype=torch.int64, device='cuda')
    ops.sample_categorical(rows, n, q, None, None, 0, d3, 1, rng_state=state3)
    assert 0.5 < float((d1 != d3).float().mean()) < 1.0
    # inside a CUDA graph: every replay advances the step
    small = rows[:300].contiguous()
    out = torch.zeros(300, dtype=torch.uint8, device='cuda')
    st = torch.tensor([99, 0, 0], dtype=torch.int64, device='cuda')
    ops.sample_categorical(small, 300, q, None, None, 0, out, 1, rng_state=st)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        ops.sample_categorical(small, 300, q, None, None, 0, out, 1, rng_state=st)
    seen = []
    for _ in range(3):
        graph.replay()
        torch.cuda.synchronize()
        seen.append(out.clone())
    assert int(st[1]) == 4 and not torch.equal(seen[0], seen[1]) and not torch.equal(seen[1], seen[2])


def test_gru_single_step_many_rows_in_one_launch(ops):
    """Generation's recurrent step: steps == 1 with more than 64 rows runs as ONE launch that walks the 64-row blocks
    (weights fetched once) and must equal the per-block launches bit for bit (GRU and LSTM, ragged last block)."""
    h = 256
    for cell, ng in ((0, 3), (1, 4)):
        for bsz in (130, 256):
            gi = rnd(bsz, ng * h, seed=cell).to(BF16)
            w_hh = rnd(ng * h, h, scale=1 / math.sqrt(h), seed=1).to(BF16)
            b_hh = rnd(ng * h, scale=0.1, seed=2)
            h0 = rnd(bsz, h, scale=0.5, seed=3)
            c0 = rnd(bsz, h, scale=0.5, seed=4)

            def run(max_rows):
                old = ops.GRU_MAX_STEP_BATCH
                ops.GRU_MAX_STEP_BATCH = max_rows
                try:
                    h_ext = torch.zeros(2, bsz, h, dtype=BF16, device='cuda')
                    h_ext[0] = h0.to(BF16)
                    hall = torch.zeros(bsz, h, dtype=BF16, device='cuda')
                    hs, cs = h0.clone(), c0.clone()
                    gates = torch.zeros(bsz, (ng + 1) * h, dtype=BF16, device='cuda')
                    if cell:
                        ops.lstm_forward(gi, w_hh, b_hh, h_ext, hall, hs, cs, gates, bsz, 1, h)
                    else:
                        ops.gru_forward(gi, w_hh, b_hh, h_ext, hall, hs, gates, bsz, 1, h)
                    return hall, hs, cs, gates, h_ext
                finally:
                    ops.GRU_MAX_STEP_BATCH = old

            ref = run(64)
            got = run(512)
            assert all(torch.equal(a, b) for a, b in zip(got, ref)), (cell, bsz)


@pytest.mark.parametrize('m,n,k,batch', [(1000, 256, 192, 1), (300, 1024, 64, 3), (129, 96, 72, 1)])
def test_gemm_relu_mask_and_gate_mask(ops, m, n, k, batch):
    """ReLU as a bit mask: the forward GEMM writes bit (row, col) = result > 0 next to the activation; a backward GEMM
    gated by those words equals one gated by the saved activation (aux_mode 2) bit for bit."""
    a = rnd(batch * m, k).to(BF16)
    w = rnd(n, k, scale=0.2, seed=1).to(BF16)
    bias = rnd(n, seed=2)
    h = torch.empty(batch * m, n, dtype=BF16, device='cuda')
    words = (n + 31) // 32
    mask = torch.zeros(batch * m, words, dtype=torch.int32, device='cuda')
    ops.gemm_nt(a, w, h, m, n, k, k, k, n, batch=batch, a_bs=m * k, c_bs=m * n, bias=bias, relu=True, relu_mask=mask)
    bits = ((mask.view(batch * m, words, 1) >> torch.arange(32, device='cuda').view(1, 1, 32)) & 1).reshape(batch * m, -1)[:, :n]
    assert torch.equal(bits.bool(), h > 0)
    g = rnd(batch * m, k, seed=3).to(BF16)                                  # any other GEMM with an (m, n) result
    d_act = torch.empty(batch * m, n, dtype=BF16, device='cuda')
    d_msk = torch.empty(batch * m, n, dtype=BF16, device='cuda')
    cs_a = torch.zeros(n, device='cuda'); cs_m = torch.zeros(n, device='cuda')
    ops.gemm_nt(g, w, d_act, m, n, k, k, k, n, batch=batch, a_bs=m * k, c_bs=m * n, aux=h, ldaux=n, aux_bs=m * n, aux_mode=2,
                colsum=cs_a)
    ops.gemm_nt(g, w, d_msk, m, n, k, k, k, n, batch=batch, a_bs=m * k, c_bs=m * n, gate_mask=mask, colsum=cs_m)
    assert torch.equal(d_act, d_msk)
    assert rel_l2(cs_m, cs_a) < 1e-5


@pytest.mark.parametrize('cell', [0, 1])
@pytest.mark.parametrize('bsz,t,h', [(128, 40, 256), (150, 9, 1024), (70, 33, 64)])
def test_recurrence_many_rows_per_launch_equals_group_launches(ops, cell, bsz, t, h):
    """More than 64 rows in ONE multi-timestep launch (groups of 64 walked inside every timestep, state of a row kept in
    its global buffer between timesteps) must equal the per-group launches bit for bit, forward and backward, GRU and
    LSTM, with a ragged last group."""
    ng = 4 if cell else 3
    gi = rnd(bsz, t, ng * h, seed=cell).to(BF16)
    w_hh = rnd(ng * h, h, scale=1 / math.sqrt(h), seed=1).to(BF16)
    b_hh = rnd(ng * h, scale=0.1, seed=2)
    h0 = rnd(bsz, h, scale=0.5, seed=3)
    c0 = rnd(bsz, h, scale=0.5, seed=4)
    dh_out = rnd(bsz, t, h, scale=0.1, seed=5).to(BF16)

    def run(max_rows):
        old = ops.GRU_MAX_STEP_BATCH
        ops.GRU_MAX_STEP_BATCH = max_rows
        try:
            h_ext = torch.zeros(t + 1, bsz, h, dtype=BF16, device='cuda')
            h_ext[0] = h0.to(BF16)
            hall = torch.zeros(bsz * t, h, dtype=BF16, device='cuda')
            hs, cs = h0.clone(), c0.clone()
            gates = torch.zeros(bsz * t, (ng + 1) * h, dtype=BF16, device='cuda')
            dgi = torch.zeros(bsz * t, ng * h, dtype=BF16, device='cuda')
            dgh = torch.zeros(t * bsz, ng * h, dtype=BF16, device='cuda')
            dh0 = torch.zeros(bsz, h, dtype=F32, device='cuda')
            dc0 = torch.zeros(bsz, h, dtype=F32, device='cuda')
            db_ih = torch.zeros(ng * h, dtype=F32, device='cuda')
            db_hh = torch.zeros(ng * h, dtype=F32, device='cuda')
            if cell:
                ops.lstm_forward(gi.view(bsz * t, ng * h), w_hh, b_hh, h_ext, hall, hs, cs, gates, bsz, t, h)
                ops.lstm_backward(w_hh.t().contiguous(), h_ext, gates, c0, dh_out.view(bsz * t, h), dgi, dgh, dh0, dc0, bsz, t, h,
                                  db_ih, db_hh)
            else:
                ops.gru_forward(gi.view(bsz * t, ng * h), w_hh, b_hh, h_ext, hall, hs, gates, bsz, t, h)
                ops.gru_backward(w_hh.t().contiguous(), h_ext, gates, dh_out.view(bsz * t, h), dgi, dgh, dh0, bsz, t, h, db_ih, db_hh)
            return hall, hs, cs, h_ext, gates, dgi, dgh, dh0, dc0, db_ih, db_hh
        finally:
            ops.GRU_MAX_STEP_BATCH = old

    ref = run(64)
    got = run(512)
    for i, (a, b) in enumerate(zip(got[:-2], ref[:-2])):
        assert torch.equal(a, b), (cell, bsz, i)
    for a, b in zip(got[-2:], ref[-2:]):                # bias gradients: same addends, summed in a different order
        assert rel_l2(a, b) < 1e-5
