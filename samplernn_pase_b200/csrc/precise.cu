// The fp32-tolerance arithmetic mode (SampleRNNModel(precision='fp32')): every activation, gate and gradient is an fp32
// tensor; the contractions still run on the tcgen05 GEMM of gemm.cu, but on SPLIT operands concatenated along K into ONE
// bf16 GEMM with fp32 accumulation: two pieces per operand and three products (a_lo.w_hi + a_hi.w_lo + a_hi.w_hi, ~4.5e-6)
// for the backward contractions, three pieces and six products (~1e-6 at K = 1024) for the forward ones.  The segments
// are ordered SMALLEST PRODUCT FIRST because the tensor core's fp32 accumulator truncates (see split3_kernel).  This file
// holds what the mode needs besides the GEMM: the split, fp32 variants of the operand-assembly kernels, the elementwise
// epilogue (aux add / ReLU / mask), an exact gather-sum for the folded embedding table, the log-softmax + NLL rows, and
// the recurrence as a host loop of (split-operand GEMM, fp32 cell kernel) per timestep.  The mode exists for parity at the
// reference's own (fp32) tolerances; it is ~5.7x slower than the bf16 path and is not what bench.py
// measures by default.
#include "common.cuh"

namespace srnn {

static inline unsigned blocks_for(long long n, int threads) { return static_cast<unsigned>((n + threads - 1) / threads); }

// Two-piece roles (x ~ hi + lo, 2^-18 relative; products a_lo.w_hi + a_hi.w_lo + a_hi.w_hi):
//   role 0: [lo | hi | hi]   (the "A" side of a product)      role 1: [hi | lo | hi]   (the "B" side)
//   role 2: [lo | hi]        (against an operand that is exact in bf16, e.g. one-hot rows: [x | x] . [lo | hi])
// Three-piece roles (x = p0 + p1 + p2 to 2^-24: fp32 exactly; the six products of order <= 2^-16):
//   role 3: [p2 | p1 | p0 | p1 | p0 | p0]   (A side)          role 4: [q0 | q1 | q2 | q0 | q1 | q0]   (B side)
// SMALLEST PRODUCTS FIRST: the tensor core's fp32 accumulator truncates (~3e-8 relative to the accumulator per K=16
// update, measured), so the order of the segments along K matters - with the leading product hi.hi LAST the accumulator
// is small (2^-8 of the result) during all the other updates and only the last K/16 updates truncate at full magnitude.
__global__ void split3_kernel(const float* __restrict__ in, long long rows, int cols, long long ld_in,
                              __nv_bfloat16* __restrict__ out, int cols_pad, long long ld_out, int role) {
  const long long g = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long r = g / cols_pad;
  if (r >= rows) return;
  const int c = static_cast<int>(g - r * cols_pad);
  const float x = c < cols ? in[r * ld_in + c] : 0.f;
  const __nv_bfloat16 h = __float2bfloat16_rn(x);
  const float r1 = x - __bfloat162float(h);
  const __nv_bfloat16 l = __float2bfloat16_rn(r1);
  __nv_bfloat16* o = out + r * ld_out + c;
  if (role == 0) {
    o[0] = l; o[cols_pad] = h; o[2 * cols_pad] = h;
  } else if (role == 1) {
    o[0] = h; o[cols_pad] = l; o[2 * cols_pad] = h;
  } else if (role == 2) {
    o[0] = l; o[cols_pad] = h;
  } else {
    const __nv_bfloat16 t = __float2bfloat16_rn(r1 - __bfloat162float(l));
    if (role == 3) {
      o[0] = t; o[cols_pad] = l; o[2 * cols_pad] = h; o[3 * cols_pad] = l; o[4 * cols_pad] = h; o[5 * cols_pad] = h;
    } else {
      o[0] = h; o[cols_pad] = l; o[2 * cols_pad] = t; o[3 * cols_pad] = h; o[4 * cols_pad] = l; o[5 * cols_pad] = h;
    }
  }
}

// out[(b, j), :] = sum_{k < r0} table[k * q + idx[b * idx_ld + j + k], :]   (fp32 rows; the embedding + conv1d + comb_layer
// embedding block as a gather-sum of the folded table, model.py:192-200 - exact in fp32, no product involved)
__global__ void embed_gather_f32_kernel(const float* __restrict__ table, const uint8_t* __restrict__ idx, long long idx_ld,
                                        int batch, int rows_per_slot, int r0, int q, int H, float* __restrict__ out) {
  const long long g = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int per_row = H / 4;
  const long long row = g / per_row;
  if (row >= static_cast<long long>(batch) * rows_per_slot) return;
  const int seg = static_cast<int>(g - row * per_row);
  const int b = static_cast<int>(row / rows_per_slot);
  const int j = static_cast<int>(row - static_cast<long long>(b) * rows_per_slot);
  const uint8_t* win = idx + b * idx_ld + j;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int k = 0; k < r0; ++k) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(table + (static_cast<long long>(k) * q + win[k]) * H) + seg);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  reinterpret_cast<float4*>(out + row * H)[seg] = acc;
}

// conditioning mixer operand (model.py:60-72), fp32
__global__ void mixer_input_f32_kernel(const float* __restrict__ utt, const float* __restrict__ table,
                                       const int* __restrict__ spk, int batch, int frames, int U, int S,
                                       float* __restrict__ out, int k_pad) {
  const long long g = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long row = g / k_pad;
  if (row >= static_cast<long long>(batch) * frames) return;
  const int c = static_cast<int>(g - row * k_pad);
  const int b = static_cast<int>(row / frames);
  float v = 0.f;
  if (c < S) v = table[static_cast<long long>(spk[b]) * S + c];
  else if (c < S + U) v = utt[row * U + (c - S)];
  out[row * k_pad + c] = v;
}

__global__ void mixer_input_bwd_f32_kernel(const float* __restrict__ d_in, const int* __restrict__ spk, int frames, int S,
                                           int k_pad, float* __restrict__ d_table) {
  const int b = blockIdx.x;
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    float acc = 0.f;
    for (int l = 0; l < frames; ++l) acc += d_in[(static_cast<long long>(b) * frames + l) * k_pad + s];
    atomicAdd(d_table + static_cast<long long>(spk[b]) * S + s, acc);
  }
}

// frame tier operand (model.py:142-147,268-271), fp32
__global__ void tier_input_f32_kernel(const uint8_t* __restrict__ xq, long long xq_ld, int x_off,
                                      const float* __restrict__ lut, const float* __restrict__ frames,
                                      const float* __restrict__ conds, int batch, int T, int fs, int L, int C,
                                      float* __restrict__ out, int k_pad) {
  const long long g = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long row = g / k_pad;
  if (row >= static_cast<long long>(batch) * T) return;
  const int c = static_cast<int>(g - row * k_pad);
  const int b = static_cast<int>(row / T);
  const int t = static_cast<int>(row - static_cast<long long>(b) * T);
  float v = 0.f;
  if (c < fs) {
    v = frames ? frames[row * fs + c] : __ldg(lut + xq[b * xq_ld + x_off + static_cast<long long>(t) * fs + c]);
  } else if (c < fs + C) {
    const int rep = T / L;
    v = conds[(static_cast<long long>(b) * L + t / rep) * C + (c - fs)];
  }
  out[row * k_pad + c] = v;
}

__global__ void tier_input_bwd_f32_kernel(const float* __restrict__ d_in, int batch, int T, int fs, int L, int C,
                                          long long ld, float* __restrict__ dconds) {
  const long long g = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (g >= static_cast<long long>(batch) * L * C) return;
  const int c = static_cast<int>(g % C);
  const long long bl = g / C;
  const int l = static_cast<int>(bl % L);
  const int b = static_cast<int>(bl / L);
  const int rep = T / L;
  float acc = 0.f;
  for (int i = 0; i < rep; ++i)
    acc += d_in[(static_cast<long long>(b) * T + static_cast<long long>(l) * rep + i) * ld + fs + c];
  dconds[g] += acc;
}

// weight-norm reparametrisation + permute into GEMM layouts, fp32 outputs (one CTA per dim-0 slice)
struct Strides3f {
  long long s[3];
};

__global__ void weight_prep_f32_kernel(const float* __restrict__ v, const float* __restrict__ g, int A, int Bd,
                                       float* __restrict__ o1, Strides3f s1, float* __restrict__ o2, Strides3f s2,
                                       float* __restrict__ inv_norm) {
  __shared__ float scratch[32];
  const int r = blockIdx.x;
  const long long n = static_cast<long long>(A) * Bd;
  const float* row = v + r * n;
  float scale = 1.f;
  if (g) {
    float ss = 0.f;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) ss += row[i] * row[i];
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = ss;
    __syncthreads();
    float t = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += scratch[w];
    const float inv = 1.f / sqrtf(t);
    if (threadIdx.x == 0 && inv_norm) inv_norm[r] = inv;
    scale = g[r] * inv;
  }
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const int a = static_cast<int>(i / Bd), b = static_cast<int>(i - static_cast<long long>(a) * Bd);
    const float w = row[i] * scale;
    if (o1) o1[r * s1.s[0] + a * s1.s[1] + b * s1.s[2]] = w;
    if (o2) o2[r * s2.s[0] + a * s2.s[1] + b * s2.s[2]] = w;
  }
}

// x[r, c] = act(x[r, c] + aux[(r / div), c] + aux2[r, c]); optional ReLU bit mask (same word layout as the GEMM's relu_mask)
__global__ void bias_act_f32_kernel(float* __restrict__ x, long long rows, int cols, long long ld,
                                    const float* __restrict__ aux, long long ldaux, int div,
                                    const float* __restrict__ aux2, long long ldaux2, int relu,
                                    uint32_t* __restrict__ mask, long long ldmask) {
  // one warp per (row, 32-column word)
  const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int words = (cols + 31) / 32;
  const long long r = warp / words;
  if (r >= rows) return;
  const int c = static_cast<int>(warp - r * words) * 32 + lane;
  float v = 0.f;
  if (c < cols) {
    v = x[r * ld + c];
    if (aux) v += aux[(r / div) * ldaux + c];
    if (aux2) v += aux2[r * ldaux2 + c];
    if (relu) v = fmaxf(v, 0.f);
    x[r * ld + c] = v;
  }
  if (mask) {
    const uint32_t m = __ballot_sync(0xffffffffu, c < cols && v > 0.f);
    if (lane == 0) mask[r * ldmask + (c >> 5)] = m;
  }
}

// din[r, c] = sum_{i < rep} dout[r * rep + i, c]
__global__ void segment_sum_f32_kernel(const float* __restrict__ dout, long long rows, int cols, long long ld_dout,
                                       int rep, float* __restrict__ din, long long ld_din) {
  const long long g = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long r = g / cols;
  if (r >= rows) return;
  const int c = static_cast<int>(g - r * cols);
  float acc = 0.f;
  for (int i = 0; i < rep; ++i) acc += dout[(r * rep + i) * ld_dout + c];
  din[r * ld_din + c] = acc;
}

// column sums of an fp32 matrix: block (32, 8), a slab of rows per block, fp32 atomics into a zeroed vector
__global__ void colsum_f32_kernel(const float* __restrict__ in, long long rows, int cols, long long ld,
                                  long long rows_per_block, float* __restrict__ out) {
  __shared__ float s[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  float acc = 0.f;
  if (c < cols)
    for (long long r = r0 + threadIdx.y; r < r1; r += 8) acc += in[r * ld + c];
  s[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float v = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) v += s[y][threadIdx.x];
    atomicAdd(out + c, v);
  }
}

// log-softmax + the target pick (model.py:203, runner.py:52), one warp per row of q <= 256 * 4 logits; the row is
// replaced by its log-probabilities
__global__ void logsoftmax_nll_f32_kernel(float* __restrict__ x, long long ld, long long m, int q,
                                          const uint8_t* __restrict__ target, float* __restrict__ lse,
                                          float* __restrict__ logp_target) {
  const long long r = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= m) return;
  float* row = x + r * ld;
  float mx = -INFINITY;
  for (int c = lane; c < q; c += 32) mx = fmaxf(mx, row[c]);
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int c = lane; c < q; c += 32) sum += expf(row[c] - mx);
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float l = mx + logf(sum);
  const int tg = target ? target[r] : -1;
  for (int c = lane; c < q; c += 32) {
    const float lp = row[c] - l;
    row[c] = lp;
    if (c == tg && logp_target) logp_target[r] = lp;
  }
  if (lane == 0 && lse) lse[r] = l;
}

// mode 0: dlogits = row_grad[r] * (onehot(target) - softmax)          (the gradient of log p(target))
// mode 1: dlogits = g - softmax * sum_c g                             (the gradient of the full log-probability row)
__global__ void logsoftmax_nll_bwd_f32_kernel(const float* __restrict__ logp, long long ld, long long m, int q,
                                              const uint8_t* __restrict__ target, const float* __restrict__ row_grad,
                                              const float* __restrict__ g, long long ldg, float* __restrict__ dl,
                                              long long lddl, int mode) {
  const long long r = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= m) return;
  const float* row = logp + r * ld;
  if (mode == 0) {
    const float rg = row_grad[r];
    const int tg = target[r];
    for (int c = lane; c < q; c += 32) dl[r * lddl + c] = rg * ((c == tg ? 1.f : 0.f) - expf(row[c]));
  } else {
    float sum = 0.f;
    for (int c = lane; c < q; c += 32) sum += g[r * ldg + c];
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    for (int c = lane; c < q; c += 32) dl[r * lddl + c] = g[r * ldg + c] - expf(row[c]) * sum;
  }
}

// ---------------------------------------------------------------------------------------------
// GRU cell (torch.nn.GRU semantics, model.py:110,152), fp32, one thread per (row, unit)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoid_acc(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ void store_split6(__nv_bfloat16* o, int kp, float x) {   // role 3: [p2 | p1 | p0 | p1 | p0 | p0]
  const __nv_bfloat16 h = __float2bfloat16_rn(x);
  const float r1 = x - __bfloat162float(h);
  const __nv_bfloat16 l = __float2bfloat16_rn(r1);
  o[0] = __float2bfloat16_rn(r1 - __bfloat162float(l));
  o[kp] = l; o[2 * kp] = h; o[3 * kp] = l; o[4 * kp] = h; o[5 * kp] = h;
}

__device__ __forceinline__ void store_split(__nv_bfloat16* o, int kp, float x) {   // role 0: [lo | hi | hi]
  const __nv_bfloat16 h = __float2bfloat16_rn(x);
  o[0] = __float2bfloat16_rn(x - __bfloat162float(h));
  o[kp] = h;
  o[2 * kp] = h;
}

// a3 <- split(h_state) before the first timestep
__global__ void gru_f32_prime_kernel(const float* __restrict__ h_state, int batch, int H, int kp,
                                     __nv_bfloat16* __restrict__ a3) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= batch * kp) return;
  const int b = g / kp, j = g - b * kp;
  store_split6(a3 + static_cast<long long>(b) * 6 * kp + j, kp, j < H ? h_state[b * H + j] : 0.f);
}

__global__ void gru_f32_cell_kernel(const float* __restrict__ gi, const float* __restrict__ gh,
                                    const float* __restrict__ b_hh, float* __restrict__ h_state,
                                    float* __restrict__ hall, float* __restrict__ gates, int batch, int T, int t, int H,
                                    int kp, __nv_bfloat16* __restrict__ a3) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= batch * H) return;
  const int b = g / H, j = g - b * H;
  const long long row = static_cast<long long>(b) * T + t;
  const float* gir = gi + row * 3 * H;
  const float* ghr = gh + static_cast<long long>(b) * 3 * H;
  const float r = sigmoid_acc(gir[j] + ghr[j] + b_hh[j]);
  const float z = sigmoid_acc(gir[H + j] + ghr[H + j] + b_hh[H + j]);
  const float hn = ghr[2 * H + j] + b_hh[2 * H + j];
  const float n = tanhf(gir[2 * H + j] + r * hn);
  const float hp = h_state[g];
  const float h = (1.f - z) * n + z * hp;
  h_state[g] = h;
  hall[row * H + j] = h;
  float* gr = gates + row * 4 * H;
  gr[j] = r;
  gr[H + j] = z;
  gr[2 * H + j] = n;
  gr[3 * H + j] = hn;
  store_split6(a3 + static_cast<long long>(b) * 6 * kp + j, kp, h);
}

// one backward timestep: dh = dh_out[t] + carry + rec; writes dgi[t], dgh[t] (+ its split operand) and the new carry dh*z
__global__ void gru_f32_cell_bwd_kernel(const float* __restrict__ gates, const float* __restrict__ hall,
                                        const float* __restrict__ h_init, const float* __restrict__ dh_out,
                                        float* __restrict__ carry, const float* __restrict__ rec,
                                        float* __restrict__ dgi, float* __restrict__ dgh, int batch, int T, int t, int H,
                                        int kp, __nv_bfloat16* __restrict__ a3) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= batch * H) return;
  const int b = g / H, j = g - b * H;
  const long long row = static_cast<long long>(b) * T + t;
  const float* gr = gates + row * 4 * H;
  const float r = gr[j], z = gr[H + j], n = gr[2 * H + j], hn = gr[3 * H + j];
  const float hp = t > 0 ? hall[(row - 1) * H + j] : h_init[g];
  const float dh = dh_out[row * H + j] + carry[g] + rec[g];
  const float dn = dh * (1.f - z) * (1.f - n * n);
  const float dz = dh * (hp - n) * z * (1.f - z);
  const float dr = dn * hn * r * (1.f - r);
  float* a = dgi + row * 3 * H;
  a[j] = dr;
  a[H + j] = dz;
  a[2 * H + j] = dn;
  float* c = dgh + row * 3 * H;
  c[j] = dr;
  c[H + j] = dz;
  c[2 * H + j] = dn * r;
  carry[g] = dh * z;
  __nv_bfloat16* o = a3 + static_cast<long long>(b) * 3 * kp;
  store_split(o + j, kp, dr);
  store_split(o + H + j, kp, dz);
  store_split(o + 2 * H + j, kp, dn * r);
}

__global__ void add2_f32_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int n) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n) out[g] = a[g] + b[g];
}

}  // namespace srnn

using namespace srnn;
#define ST(s) static_cast<cudaStream_t>(s)

extern "C" int srnn_split3_bf16(const float* in, int64_t rows, int32_t cols, int64_t ld_in, void* out, int32_t cols_pad,
                                int64_t ld_out, int32_t role, srnn_stream_t s) {
  SRNN_CHECK_ARG(in && out && rows > 0 && cols > 0 && cols_pad >= cols && cols_pad % 8 == 0 && role >= 0 && role <= 4,
                 "split3_bf16: cols_pad must be a multiple of 8 and >= cols, role in 0..4");
  SRNN_CHECK_ARG(ld_out >= (role == 2 ? 2 : role >= 3 ? 6 : 3) * static_cast<int64_t>(cols_pad),
                 "split3_bf16: ld_out too small");
  split3_kernel<<<blocks_for(rows * cols_pad, 256), 256, 0, ST(s)>>>(in, rows, cols, ld_in,
                                                                     static_cast<__nv_bfloat16*>(out), cols_pad, ld_out, role);
  SRNN_CUDA(cudaGetLastError());
  return SRNN_OK;
}

extern "C" int srnn_mixer_input_f32(const float* utt, const float* table, const int32_t* spk, int32_t batch,
                                    int32_t frames, int32_t U, int32_t S, float* out, int32_t k_pad, srnn_stream_t s) {
  SRNN_CHECK_ARG(utt && table && spk && out && batch > 0 && frames > 0 && k_pad >= U + S, "mixer_input_f32: bad arguments");
  mixer_input_f32_kernel<<<blocks_for(static_cast<long long>(batch) * frames * k_pad, 256), 256, 0, ST(s)>>>(
      utt, table, spk, batch, frames, U, S, out, k_pad);
  SRNN_CUDA(cudaGetLastError());
  return SRNN_OK;
}

extern "C" int srnn_mixer_input_bwd_f32(const float* d_in, const int32_t* spk, int32_t batch, int32_t frames, int32_t S,
                                        int32_t k_pad, float* d_table, srnn_stream_t s) {
  SRNN_CHECK_ARG(d_in && spk && d_table && batch > 0 && frames > 0 && S > 0, "mixer_input_bwd_f32: bad arguments");
  mixer_input_bwd_f32_kernel<<<batch, 128, 0, ST(s)>>>(d_in, spk, frames, S, k_pad, d_table);
  SRNN_CUDA(cudaGetLastError());
  return SRNN_OK;
}

extern "C" int srnn_tier_input_f32(const uint8_t* xq, int64_t xq_ld, int32_t x_off, const float* lut,
                                   const float* frames, const float* conds, int32_t batch, int32_t T, int32_t fs,
                                   int32_t L, int32_t C, float* out, int32_t k_pad, srnn_stream_t s) {
  SRNN_CHECK_ARG(((xq && lut) || frames) && conds && out && batch > 0 && T > 0 && L > 0 && T % L == 0 && k_pad >= fs + C,
                 "tier_input_f32: bad arguments");
  tier_input_f32_kernel<<<blocks_for(static_cast<long long>(batch) * T * k_pad, 256), 256, 0, ST(s)>>>(
      xq, xq_ld, x_off, lut, frames, conds, batch, T, fs, L, C, out, k_pad);
  SRNN_CUDA(cudaGetLastError());
  return SRNN_OK;
}

extern "C" int srnn_tier_input_bwd_f32(const float* d_in, int32_t batch, int32_t T, int32_t fs, int32_t L, int32_t C,
                                       int64_t ld, float* dconds, srnn_stream_t s) {
  SRNN_CHECK_ARG(d_in && dconds && batch > 0 && T > 0 && L > 0 && T % L == 0, "tier_input_bwd_f32: bad arguments");
  tier_input_bwd_f32_kernel<<<blocks_for(static_cast<long long>(batch) * L * C, 256), 256, 0, ST(s)>>>(
      d_in, batch, T, fs, L, C, ld, dconds);
  SRNN_CUDA(cudaGetLastError());
  return SRNN_OK;
}

extern "C" int srnn_weight_prep_f32(const float* v, const float* g, int32_t R, int32_t A, int32_t B, float* o1,
                                    const int64_t* s1, float* o2, const int64_t* s2, float* inv_norm, srnn_stream_t s) {
  SRNN_CHECK_ARG(v && R > 0 && A > 0 && B > 0 && o1 && s1, "weight_prep_f32: bad arguments");
  SRNN_CHECK_ARG(!o2 || s2, "weight_prep_f32: second output needs strides");
  Strides3f a{{s1[0], s1[1], s1[2]}}, b{{0, 0, 0}};
  if (o2) b = Strides3f{{s2[0], s2[1], s2[2]}};
  weight_prep_f32_kernel<<<R, 256, 0, ST(s)>>>(v, g, A, B, o1, a, o2, b, inv_norm);
  SRNN_CUDA(cudaGetLastError());
  return SRNN_OK;
}

extern "C" int srnn_bias_act_f32(float* x, int64_t rows, int32_t cols, int64_t ld, const float* aux, int64_t ldaux,
                                 int32_t aux_row_div, const float* aux2, int64_t ldaux2, int32_t relu, uint32_t* mask,
                                 int64_t ldmask, srnn_stream_t s) {
  SRNN_CHECK_ARG(x && rows > 0 && cols > 0, "bias_act_f32: bad arguments");
  const long long warps = rows * ((cols + 31) / 32);
  bias_act_f32_kernel<<<blocks_for(warps * 32, 256), 256, 0, ST(s)>>>(x, rows, cols, ld, aux, ldaux,
                                                                      aux_row_div > 0 ? aux_row_div : 1, aux2, ldaux2, relu,
                                                                      mask, ldmask);
  SRNN_CUDA(cudaGetLastError());
  return SRNN_OK;
}

extern "C" int srnn_segment_sum_f32(const float* dout, int64_t rows, int32_t cols, int64_t ld_dout, int32_t rep,
                                    float* din, int64_t ld_din, srnn_stream_t s) {
  SRNN_CHECK_ARG(dout && din && rows > 0 && cols > 0 && rep > 0, "segment_sum_f32: bad arguments");
  segment_sum_f32_kernel<<<blocks_for(rows * cols, 256), 256, 0, ST(s)>>>(dout, rows, cols, ld_dout, rep, din, ld_din);
  SRNN_CUDA(cudaGetLastError());
  return SRNN_OK;
}

extern "C" int srnn_colsum_f32(const float* in, int64_t rows, int32_t cols, int64_t ld, float* out, srnn_stream_t s) {
  SRNN_CHECK_ARG(in && out && rows > 0 && cols > 0, "colsum_f32: bad arguments");
  SRNN_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * cols, ST(s)));
  const int col_blocks = (cols + 31) / 32;
  long long slabs = (8LL * sm_count() + col_blocks - 1) / col_blocks;
  if (slabs > (rows + 63) / 64) slabs = (rows + 63) / 64;
  if (slabs < 1) slabs = 1;
  if (slabs > 65535) slabs = 65535;
  const long long rpb = (rows + slabs - 1) / slabs;
  colsum_f32_kernel<<<dim3(col_blocks, static_cast<unsigned>(slabs)), dim3(32, 8), 0, ST(s)>>>(in, rows, cols, ld, rpb,
                                                                                                out);
  SRNN_CUDA(cudaGetLastError());
  return SRNN_OK;
}

extern "C" int srnn_logsoftmax_nll_f32(float* x, int64_t ld, int64_t m, int32_t q, const uint8_t* target, float* lse,
                                       float* logp_target, srnn_stream_t s) {
  SRNN_CHECK_ARG(x && m > 0 && q > 0 && q <= 256 && (!logp_target || target), "logsoftmax_nll_f32: bad arguments");
  logsoftmax_nll_f32_kernel<<<blocks_for(m * 32, 256), 256, 0, ST(s)>>>(x, ld, m, q, target, lse, logp_target);
  SRNN_CUDA(cudaGetLastError());
  return SRNN_OK;
}

extern "C" int srnn_logsoftmax_nll_bwd_f32(const float* logp, int64_t ld, int64_t m, int32_t q, const uint8_t* target,
                                           const float* row_grad, const float* g, int64_t ldg, float* dlogits,
                                           int64_t lddl, srnn_stream_t s) {
  SRNN_CHECK_ARG(logp && dlogits && m > 0 && q > 0 && ((row_grad && target) || g),
                 "logsoftmax_nll_bwd_f32: needs (row_grad, target) or g");
  logsoftmax_nll_bwd_f32_kernel<<<blocks_for(m * 32, 256), 256, 0, ST(s)>>>(logp, ld, m, q, target, row_grad, g, ldg,
                                                                            dlogits, lddl, g ? 1 : 0);
  SRNN_CUDA(cudaGetLastError());
  return SRNN_OK;
}

extern "C" int srnn_embed_gather_f32(const float* table, const uint8_t* idx, int64_t idx_ld, int32_t batch,
                                     int32_t rows_per_slot, int32_t r0, int32_t q, int32_t hidden, float* out,
                                     srnn_stream_t s) {
  SRNN_CHECK_ARG(table && idx && out && batch > 0 && rows_per_slot > 0 && r0 > 0 && q > 0 && hidden % 4 == 0,
                 "embed_gather_f32: hidden must be a multiple of 4");
  embed_gather_f32_kernel<<<blocks_for(static_cast<long long>(batch) * rows_per_slot * (hidden / 4), 256), 256, 0, ST(s)>>>(
      table, idx, idx_ld, batch, rows_per_slot, r0, q, hidden, out);
  SRNN_CUDA(cudaGetLastError());
  return SRNN_OK;
}

// one recurrent product of the fp32 mode: C[batch, n] = A[batch, segs kp] . W[n, segs kp]^T on the tcgen05 GEMM
static int split_gemm(const void* a3, const void* w3, float* c, int batch, int n, int kp, int segs, srnn_stream_t s) {
  srnn_gemm_args g{};
  g.op = 0;
  g.m = batch; g.n = n; g.k = segs * kp; g.batch = 1;
  g.a = a3; g.lda = segs * kp;
  g.b = w3; g.ldb = segs * kp;
  g.c = c; g.ldc = n; g.c_dtype = 1;
  g.aux_row_div = 1;
  return srnn_gemm_bf16(&g, s);
}

extern "C" int srnn_gru_forward_f32(const srnn_gru_f32_args* a, srnn_stream_t s) {
  SRNN_CHECK_ARG(a && a->batch > 0 && a->steps > 0 && a->hidden > 0 && a->hidden % 8 == 0, "gru_forward_f32: bad sizes");
  SRNN_CHECK_ARG(a->gi && a->w3 && a->b_hh && a->h_state && a->hall && a->gates && a->a3 && a->ws,
                 "gru_forward_f32: null buffer");
  const int B = a->batch, T = a->steps, H = a->hidden, kp = H;
  gru_f32_prime_kernel<<<blocks_for(static_cast<long long>(B) * kp, 256), 256, 0, ST(s)>>>(
      a->h_state, B, H, kp, static_cast<__nv_bfloat16*>(a->a3));
  SRNN_CUDA(cudaGetLastError());
  for (int t = 0; t < T; ++t) {
    const int rc = split_gemm(a->a3, a->w3, a->ws, B, 3 * H, kp, 6, s);
    if (rc) return rc;
    gru_f32_cell_kernel<<<blocks_for(static_cast<long long>(B) * H, 256), 256, 0, ST(s)>>>(
        a->gi, a->ws, a->b_hh, a->h_state, a->hall, a->gates, B, T, t, H, kp, static_cast<__nv_bfloat16*>(a->a3));
    SRNN_CUDA(cudaGetLastError());
  }
  return SRNN_OK;
}

extern "C" int srnn_gru_backward_f32(const srnn_gru_f32_args* a, srnn_stream_t s) {
  SRNN_CHECK_ARG(a && a->batch > 0 && a->steps > 0 && a->hidden > 0 && a->hidden % 8 == 0, "gru_backward_f32: bad sizes");
  SRNN_CHECK_ARG(a->w3 && a->hall && a->h_init && a->gates && a->a3 && a->ws && a->dh_out && a->dgi && a->dgh && a->dh0 &&
                     a->carry,
                 "gru_backward_f32: null buffer");
  const int B = a->batch, T = a->steps, H = a->hidden, kp = 3 * H;
  SRNN_CUDA(cudaMemsetAsync(a->carry, 0, sizeof(float) * B * H, ST(s)));
  SRNN_CUDA(cudaMemsetAsync(a->ws, 0, sizeof(float) * B * H, ST(s)));
  for (int t = T - 1; t >= 0; --t) {
    gru_f32_cell_bwd_kernel<<<blocks_for(static_cast<long long>(B) * H, 256), 256, 0, ST(s)>>>(
        a->gates, a->hall, a->h_init, a->dh_out, a->carry, a->ws, a->dgi, a->dgh, B, T, t, H, kp,
        static_cast<__nv_bfloat16*>(a->a3));
    SRNN_CUDA(cudaGetLastError());
    const int rc = split_gemm(a->a3, a->w3, a->ws, B, H, kp, 3, s);  // rec[b, i] = sum_j dgh[b, j] W_hh[j, i]
    if (rc) return rc;
  }
  add2_f32_kernel<<<blocks_for(static_cast<long long>(B) * H, 256), 256, 0, ST(s)>>>(a->carry, a->ws, a->dh0, B * H);
  SRNN_CUDA(cudaGetLastError());
  return SRNN_OK;
}
