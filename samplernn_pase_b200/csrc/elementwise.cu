// HBM-bound kernels of the training step: quantiser, one-hot/operand assembly, weight-norm
// reparametrisation, reductions, state selection, loss reduction and the fused AdamClipped.
// All are single-pass, coalesced, vectorised where the layout allows; grids are sized from the
// element count (these tensors are far larger than 148 SMs x resident CTAs).
#include <math.h>

#include "common.cuh"

namespace srnn {

static inline unsigned blocks_for(long long n, int threads) {
  long long b = (n + threads - 1) / threads;
  if (b < 1) b = 1;
  return static_cast<unsigned>(b);
}

// ---------------------------------------------------------------------------------------------
// quantize_ulaw (utils.py:59-65).  The reference evaluates the chain as separate torch ops, each
// rounding to fp32; on CUDA torch turns "/ LOG_MU1" into a multiply by the fp32 reciprocal and the
// Python double (256 - 1e-6) becomes 256.0f.  __fmul_rn/__fadd_rn keep ptxas from contracting
// into FMAs, logf is the same libdevice routine torch's log kernel calls.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ long long ulaw_index(float x) {
  const float inv_log_mu1 = 1.0f / 5.5451774444795623f;
  const float a = fabsf(x);
  const float sgn = (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f);
  const float mag = logf(__fadd_rn(__fmul_rn(255.0f, a), 1.0f));
  const float s = __fmul_rn(__fmul_rn(sgn, mag), inv_log_mu1);
  const float y = __fmul_rn(0.5f, __fadd_rn(s, 1.0f));
  return static_cast<long long>(__fmul_rn(y, 256.0f));
}

__global__ void quantize_ulaw_kernel(const float* __restrict__ x, long long n, long long* __restrict__ o64,
                                     uint8_t* __restrict__ o8, int* __restrict__ overflow) {
  const long long i4 = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i4 >= n) return;
  float v[4];
  const bool vec = (i4 + 4 <= n) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  if (vec) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(x + i4));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
    for (int j = 0; j < 4; ++j) v[j] = (i4 + j < n) ? x[i4 + j] : 0.f;
  }
  long long q[4];
  int over = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    q[j] = ulaw_index(v[j]);
    over += (q[j] > 255 || q[j] < 0) ? 1 : 0;
  }
  if (over && overflow) atomicAdd(overflow, over);
  if (o64) {
    if (vec && ((reinterpret_cast<uintptr_t>(o64) & 15) == 0)) {
      reinterpret_cast<longlong2*>(o64 + i4)[0] = make_longlong2(q[0], q[1]);
      reinterpret_cast<longlong2*>(o64 + i4)[1] = make_longlong2(q[2], q[3]);
    } else {
      for (int j = 0; j < 4; ++j)
        if (i4 + j < n) o64[i4 + j] = q[j];
    }
  }
  if (o8) {
    uint8_t b[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = static_cast<uint8_t>(q[j] < 0 ? 0 : (q[j] > 255 ? 255 : q[j]));
    if (vec && ((reinterpret_cast<uintptr_t>(o8) & 3) == 0)) {
      *reinterpret_cast<uchar4*>(o8 + i4) = make_uchar4(b[0], b[1], b[2], b[3]);
    } else {
      for (int j = 0; j < 4; ++j)
        if (i4 + j < n) o8[i4 + j] = b[j];
    }
  }
}

// quantize_linear (utils.py:48-54) with per-row min/max: one CTA per row.
__global__ void quantize_linear_kernel(const float* __restrict__ x, long long cols, float scale, long long* __restrict__ o64,
                                       uint8_t* __restrict__ o8) {
  __shared__ float smin[32], smax[32];
  const float* row = x + blockIdx.x * cols;
  float lo = INFINITY, hi = -INFINITY;
  for (long long i = threadIdx.x; i < cols; i += blockDim.x) {
    lo = fminf(lo, row[i]);
    hi = fmaxf(hi, row[i]);
  }
  for (int o = 16; o; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    smin[threadIdx.x >> 5] = lo;
    smax[threadIdx.x >> 5] = hi;
  }
  __syncthreads();
  lo = smin[0];
  hi = smax[0];
  for (int w = 1; w < (blockDim.x >> 5); ++w) {
    lo = fminf(lo, smin[w]);
    hi = fmaxf(hi, smax[w]);
  }
  const float range = __fsub_rn(hi, lo);   // max of (x - min) == max - min exactly (monotone rounding)
  for (long long i = threadIdx.x; i < cols; i += blockDim.x) {
    float y = __fsub_rn(row[i], lo);
    y = __fdiv_rn(y, range);
    y = __fmul_rn(y, scale);                // q_levels - 1e-2, rounded to fp32 like the reference's in-place multiply
    y = __fadd_rn(y, 0.005f);
    // a constant row is 0/0 in the reference (NaN -> an undefined integer); here it maps to index 0
    const long long q = range > 0.f ? static_cast<long long>(y) : 0ll;
    if (o64) o64[blockIdx.x * cols + i] = q;
    if (o8) o8[blockIdx.x * cols + i] = static_cast<uint8_t>(q < 0 ? 0 : (q > 255 ? 255 : q));
  }
}

__global__ void dequant_lut_kernel(const long long* __restrict__ i64, const uint8_t* __restrict__ i8, long long n,
                                   const float* __restrict__ lut, float* __restrict__ of, __nv_bfloat16* __restrict__ ob) {
  __shared__ float s[257];
  for (int i = threadIdx.x; i < 257; i += blockDim.x) s[i] = lut[i];
  __syncthreads();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long q = i64 ? i64[i] : static_cast<long long>(i8[i]);
  q = q < 0 ? 0 : (q > 256 ? 256 : q);
  const float v = s[q];
  if (of) of[i] = v;
  if (ob) ob[i] = __float2bfloat16_rn(v);
}

// one-hot rows: 32 threads per row, 16 bytes (8 bf16) each
__global__ void onehot_kernel(const uint8_t* __restrict__ idx, long long n, int q, __nv_bfloat16* __restrict__ out) {
  const int per_row = q / 8;
  const long long g = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long row = g / per_row;
  if (row >= n) return;
  const int seg = static_cast<int>(g - row * per_row);
  const int hot = idx[row];
  uint4 v = make_uint4(0, 0, 0, 0);
  if (hot >= seg * 8 && hot < seg * 8 + 8) {
    const int e = hot - seg * 8;
    const uint32_t one = 0x3F80u << ((e & 1) * 16);
    (&v.x)[e >> 1] = one;
  }
  reinterpret_cast<uint4*>(out + row * q)[seg] = v;
}

// ---------------------------------------------------------------------------------------------
// weight-norm reparametrisation + cast/permute into GEMM layouts (one CTA per dim-0 slice)
// ---------------------------------------------------------------------------------------------
struct Strides3 {
  long long s[3];
};

__device__ __forceinline__ float block_sum(float v, float* scratch) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < (blockDim.x >> 5); ++w) t += scratch[w];
  return t;
}

__global__ void weight_prep_kernel(const float* __restrict__ v, const float* __restrict__ g, int A, int Bd,
                                   __nv_bfloat16* __restrict__ o1, Strides3 s1, __nv_bfloat16* __restrict__ o2,
                                   Strides3 s2, float* __restrict__ inv_norm) {
  __shared__ float scratch[32];
  const int r = blockIdx.x;
  const long long n = static_cast<long long>(A) * Bd;
  const float* row = v + r * n;
  float scale = 1.f;
  if (g) {
    float ss = 0.f;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) ss += row[i] * row[i];
    ss = block_sum(ss, scratch);
    const float inv = 1.f / sqrtf(ss);
    if (threadIdx.x == 0 && inv_norm) inv_norm[r] = inv;
    scale = g[r] * inv;
  }
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const int a = static_cast<int>(i / Bd), b = static_cast<int>(i - static_cast<long long>(a) * Bd);
    const __nv_bfloat16 w = __float2bfloat16_rn(row[i] * scale);
    if (o1) o1[r * s1.s[0] + a * s1.s[1] + b * s1.s[2]] = w;
    if (o2) o2[r * s2.s[0] + a * s2.s[1] + b * s2.s[2]] = w;
  }
}

__global__ void weight_prep_bwd_kernel(const float* __restrict__ dw, Strides3 s, const float* __restrict__ v,
                                       const float* __restrict__ g, const float* __restrict__ inv_norm, int A, int Bd,
                                       float* __restrict__ dv, float* __restrict__ dg) {
  __shared__ float scratch[32];
  const int r = blockIdx.x;
  const long long n = static_cast<long long>(A) * Bd;
  const float* row = v + r * n;
  if (!g) {
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
      const int a = static_cast<int>(i / Bd), b = static_cast<int>(i - static_cast<long long>(a) * Bd);
      dv[r * n + i] = dw[r * s.s[0] + a * s.s[1] + b * s.s[2]];
    }
    return;
  }
  float dot = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const int a = static_cast<int>(i / Bd), b = static_cast<int>(i - static_cast<long long>(a) * Bd);
    dot += dw[r * s.s[0] + a * s.s[1] + b * s.s[2]] * row[i];
  }
  dot = block_sum(dot, scratch);
  const float inv = inv_norm[r];
  const float scale = g[r] * inv;
  if (threadIdx.x == 0) dg[r] = dot * inv;
  const float proj = dot * inv * inv;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const int a = static_cast<int>(i / Bd), b = static_cast<int>(i - static_cast<long long>(a) * Bd);
    dv[r * n + i] = scale * (dw[r * s.s[0] + a * s.s[1] + b * s.s[2]] - row[i] * proj);
  }
}

__global__ void pad_cast_kernel(const float* __restrict__ in, long long rows, int cols, long long ld_in,
                                __nv_bfloat16* __restrict__ out, int cols_pad, long long ld_out) {
  const long long g = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long r = g / cols_pad;
  if (r >= rows) return;
  const int c = static_cast<int>(g - r * cols_pad);
  out[r * ld_out + c] = __float2bfloat16_rn(c < cols ? in[r * ld_in + c] : 0.f);
}

// x = hi + lo with hi = bf16(x), lo = bf16(x - hi): two bf16 terms carry 16 mantissa bits of an fp32 value
__global__ void split_bf16_kernel(const float* __restrict__ in, long long rows, int cols, long long ld_in,
                                  __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int cols_pad,
                                  long long ld_out) {
  const long long g = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long r = g / cols_pad;
  if (r >= rows) return;
  const int c = static_cast<int>(g - r * cols_pad);
  const float x = c < cols ? in[r * ld_in + c] : 0.f;
  const __nv_bfloat16 h = __float2bfloat16_rn(x);
  hi[r * ld_out + c] = h;
  lo[r * ld_out + c] = __float2bfloat16_rn(x - __bfloat162float(h));
}

__global__ void bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ in, long long rows, int cols, long long ld_in,
                                   float* __restrict__ out, long long ld_out, int accumulate) {
  const long long g = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long r = g / cols;
  if (r >= rows) return;
  const int c = static_cast<int>(g - r * cols);
  const float v = __bfloat162float(in[r * ld_in + c]);
  float* o = out + r * ld_out + c;
  *o = accumulate ? *o + v : v;
}

// ---------------------------------------------------------------------------------------------
// conditioning mixer operand (model.py:60-72)
// ---------------------------------------------------------------------------------------------
__global__ void mixer_input_kernel(const float* __restrict__ utt, const float* __restrict__ table,
                                   const int* __restrict__ spk, int batch, int frames, int U, int S,
                                   __nv_bfloat16* __restrict__ out, int k_pad) {
  const long long g = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long row = g / k_pad;
  if (row >= static_cast<long long>(batch) * frames) return;
  const int c = static_cast<int>(g - row * k_pad);
  const int b = static_cast<int>(row / frames);
  float v = 0.f;
  if (c < S) v = table[static_cast<long long>(spk[b]) * S + c];
  else if (c < S + U) v = utt[row * U + (c - S)];
  out[row * k_pad + c] = __float2bfloat16_rn(v);
}

__global__ void mixer_input_bwd_kernel(const __nv_bfloat16* __restrict__ d_in, const int* __restrict__ spk, int frames,
                                       int S, int k_pad, float* __restrict__ d_table) {
  const int b = blockIdx.x;
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    float acc = 0.f;
    for (int l = 0; l < frames; ++l) acc += __bfloat162float(d_in[(static_cast<long long>(b) * frames + l) * k_pad + s]);
    atomicAdd(d_table + static_cast<long long>(spk[b]) * S + s, acc);
  }
}

// ---------------------------------------------------------------------------------------------
// frame tier operand (model.py:142-147,268-271)
// ---------------------------------------------------------------------------------------------
__global__ void tier_input_kernel(const uint8_t* __restrict__ xq, long long xq_ld, int x_off,
                                  const float* __restrict__ lut, const float* __restrict__ frames,
                                  const float* __restrict__ conds, int batch, int T,
                                  int fs, int L, int C, __nv_bfloat16* __restrict__ out, int k_pad) {
  __shared__ float s[256];
  if (lut)
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s[i] = lut[i];
  __syncthreads();
  const long long g = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long row = g / k_pad;
  if (row >= static_cast<long long>(batch) * T) return;
  const int c = static_cast<int>(g - row * k_pad);
  const int b = static_cast<int>(row / T);
  const int t = static_cast<int>(row - static_cast<long long>(b) * T);
  float v = 0.f;
  if (c < fs) {
    v = frames ? frames[row * fs + c] : s[xq[b * xq_ld + x_off + static_cast<long long>(t) * fs + c]];
  } else if (c < fs + C) {
    const int rep = T / L;
    v = conds[(static_cast<long long>(b) * L + t / rep) * C + (c - fs)];
  }
  out[row * k_pad + c] = __float2bfloat16_rn(v);
}

__global__ void tier_input_bwd_kernel(const __nv_bfloat16* __restrict__ d_in, int batch, int T, int fs, int L, int C,
                                      int k_pad, float* __restrict__ dconds) {
  const long long g = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (g >= static_cast<long long>(batch) * L * C) return;
  const int c = static_cast<int>(g % C);
  const long long bl = g / C;
  const int l = static_cast<int>(bl % L);
  const int b = static_cast<int>(bl / L);
  const int rep = T / L;
  float acc = 0.f;
  for (int i = 0; i < rep; ++i)
    acc += __bfloat162float(d_in[(static_cast<long long>(b) * T + static_cast<long long>(l) * rep + i) * k_pad + fs + c]);
  dconds[g] += acc;
}

// row repeat and its adjoint; 8 bf16 per thread
__global__ void repeat_rows_kernel(const __nv_bfloat16* __restrict__ in, long long rows, int cols, long long ld_in,
                                   int rep, __nv_bfloat16* __restrict__ out, long long ld_out) {
  const int per_row = cols / 8;
  const long long g = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long orow = g / per_row;
  if (orow >= rows * rep) return;
  const int seg = static_cast<int>(g - orow * per_row);
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + (orow / rep) * ld_in) + seg);
  reinterpret_cast<uint4*>(out + orow * ld_out)[seg] = v;
}

__global__ void repeat_rows_bwd_kernel(const __nv_bfloat16* __restrict__ dout, long long rows, int cols,
                                       long long ld_dout, int rep, __nv_bfloat16* __restrict__ din, long long ld_din) {
  const int per_row = cols / 8;
  const long long g = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long irow = g / per_row;
  if (irow >= rows) return;
  const int seg = static_cast<int>(g - irow * per_row);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < rep; ++i) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(dout + (irow * rep + i) * ld_dout) + seg);
    acc[0] += bf16_lo(u.x); acc[1] += bf16_hi(u.x);
    acc[2] += bf16_lo(u.y); acc[3] += bf16_hi(u.y);
    acc[4] += bf16_lo(u.z); acc[5] += bf16_hi(u.z);
    acc[6] += bf16_lo(u.w); acc[7] += bf16_hi(u.w);
  }
  uint4 o;
  o.x = pack_bf16x2(acc[0], acc[1]);
  o.y = pack_bf16x2(acc[2], acc[3]);
  o.z = pack_bf16x2(acc[4], acc[5]);
  o.w = pack_bf16x2(acc[6], acc[7]);
  reinterpret_cast<uint4*>(din + irow * ld_din)[seg] = o;
}

// column sums: block (32, 8) covers 256 columns x one row slab; every thread keeps 4 independent
// 16-byte loads in flight; fp32 atomics into a zeroed vector
__global__ void __launch_bounds__(256)
colsum_kernel(const __nv_bfloat16* __restrict__ in, long long rows, int cols, long long ld, long long rows_per_block,
              float* __restrict__ out) {
  __shared__ float s[8][257];
  const int c = blockIdx.x * 256 + threadIdx.x * 8;
  const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const bool vec = (c + 8 <= cols) && ((ld & 7) == 0) && ((reinterpret_cast<uintptr_t>(in) & 15) == 0);
  if (vec) {
    for (long long r = r0 + threadIdx.y; r < r1; r += 32) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long rr = r + u * 8;
        v[u] = rr < r1 ? __ldg(reinterpret_cast<const uint4*>(in + rr * ld + c)) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        acc[0] += bf16_lo(v[u].x); acc[1] += bf16_hi(v[u].x);
        acc[2] += bf16_lo(v[u].y); acc[3] += bf16_hi(v[u].y);
        acc[4] += bf16_lo(v[u].z); acc[5] += bf16_hi(v[u].z);
        acc[6] += bf16_lo(v[u].w); acc[7] += bf16_hi(v[u].w);
      }
    }
  } else if (c < cols) {
    for (long long r = r0 + threadIdx.y; r < r1; r += 8)
      for (int j = 0; j < 8; ++j)
        if (c + j < cols) acc[j] += __bfloat162float(in[r * ld + c + j]);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) s[threadIdx.y][threadIdx.x * 8 + j] = acc[j];
  __syncthreads();
  const int t = threadIdx.y * 32 + threadIdx.x;      // 256 threads -> 256 columns
  const int cc = blockIdx.x * 256 + t;
  if (cc < cols) {
    float v = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) v += s[y][t];
    atomicAdd(out + cc, v);
  }
}

// ---------------------------------------------------------------------------------------------
// recurrent state selection (model.py:149-151,239-243) and its adjoint
// ---------------------------------------------------------------------------------------------
__global__ void state_select_kernel(const float* __restrict__ carried, const float* __restrict__ h0,
                                    const uint8_t* __restrict__ use_carry, int batch, int H, float* __restrict__ h_state,
                                    __nv_bfloat16* __restrict__ h_ext, long long ext_ld) {
  const long long g = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (g >= static_cast<long long>(batch) * H) return;
  const int b = static_cast<int>(g / H), j = static_cast<int>(g - static_cast<long long>(b) * H);
  const float v = (use_carry[b] && carried) ? carried[g] : h0[j];
  h_state[g] = v;
  if (h_ext) h_ext[b * ext_ld + j] = __float2bfloat16_rn(v);
}

__global__ void state_select_bwd_kernel(const float* __restrict__ dh, const uint8_t* __restrict__ use_carry, int batch,
                                        int H, float* __restrict__ d_h0) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= H) return;
  float acc = 0.f;
  for (int b = 0; b < batch; ++b)
    if (!use_carry[b]) acc += dh[static_cast<long long>(b) * H + j];
  d_h0[j] = acc;
}

// ---------------------------------------------------------------------------------------------
// loss reduction (runner.py:52 over the rows kept by model.py:283-284)
// ---------------------------------------------------------------------------------------------
__global__ void masked_sum_kernel(const float* __restrict__ lp, const uint8_t* __restrict__ valid, long long rows,
                                  int rows_per_slot, float* __restrict__ out2) {
  __shared__ float scratch[32];
  float acc = 0.f, cnt = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < rows;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    if (valid[i / rows_per_slot]) {
      acc += lp[i];
      cnt += 1.f;
    }
  }
  acc = block_sum(acc, scratch);
  cnt = block_sum(cnt, scratch);
  if (threadIdx.x == 0) {
    atomicAdd(out2, acc);
    atomicAdd(out2 + 1, cnt);
  }
}
__global__ void nll_finalize_kernel(float* out2) { out2[0] = -out2[0] / out2[1]; }

// ---------------------------------------------------------------------------------------------
// AdamClipped (optimizer.py:6-14 + torch.optim.Adam defaults)
// ---------------------------------------------------------------------------------------------
__global__ void adam_clipped_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m,
                                    float* __restrict__ v, long long n, float lr, float b1, float b2, float eps,
                                    float bc1, float bc2_sqrt, float gscale) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float gi = g[i] * gscale;
  gi = fminf(fmaxf(gi, -1.f), 1.f);
  const float mi = b1 * m[i] + (1.f - b1) * gi;
  const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  w[i] -= (lr / bc1) * (mi / denom);
}

}  // namespace srnn

using namespace srnn;
#define ST(s) static_cast<cudaStream_t>(s)

extern "C" int srnn_quantize_ulaw(const float* x, int64_t n, int64_t* o64, uint8_t* o8, int32_t* overflow,
                                  srnn_stream_t s) {
  SRNN_CHECK_ARG(x && n >= 0 && (o64 || o8), "quantize_ulaw: null input/output");
  if (n == 0) return SRNN_OK;
  quantize_ulaw_kernel<<<blocks_for((n + 3) / 4, 256), 256, 0, ST(s)>>>(x, n, reinterpret_cast<long long*>(o64), o8,
                                                                      overflow);
  SRNN_CUDA(cudaGetLastError());
  return SRNN_OK;
}

extern "C" int srnn_quantize_linear(const float* x, int64_t rows, int64_t cols, int32_t q_levels, int64_t* o64, uint8_t* o8,
                                    srnn_stream_t s) {
  SRNN_CHECK_ARG(x && rows > 0 && cols > 0 && (o64 || o8), "quantize_linear: bad arguments");
  SRNN_CHECK_ARG(q_levels >= 2 && (o8 == nullptr || q_levels <= 256),
                 "quantize_linear: q_levels must be >= 2 (and <= 256 when uint8 indices are requested), got %d", q_levels);
  const float scale = static_cast<float>(static_cast<double>(q_levels) - 1e-2);
  quantize_linear_kernel<<<static_cast<unsigned>(rows), 256, 0, ST(s)>>>(x, cols, scale, reinterpret_cast<long long*>(o64), o8);
  SRNN_CUDA(cudaGetLastError());
  return SRNN_OK;
}

extern "C" int srnn_dequantize_lut(const int64_t* i64, const uint8_t* i8, int64_t n, const float* lut, float* of,
                                   void* ob, srnn_stream_t s) {
  SRNN_CHECK_ARG((i64 != nullptr) != (i8 != nullptr), "dequantize_lut: exactly one index input");
  SRNN_CHECK_ARG(lut && (of || ob), "dequantize_lut: null lut/output");
  if (n == 0) return SRNN_OK;
  dequant_lut_kernel<<<blocks_for(n, 256), 256, 0, ST(s)>>>(reinterpret_cast<const long long*>(i64), i8, n, lut, of,
                                                           static_cast<__nv_bfloat16*>(ob));
  SRNN_CUDA(cudaGetLastError());
  return SRNN_OK;
}

extern "C" int srnn_onehot_rows(const uint8_t* idx, int64_t n, int32_t q, void* out, srnn_stream_t s) {
  SRNN_CHECK_ARG(idx && out && q > 0 && q % 8 == 0 && q <= 256, "onehot_rows: q must be a multiple of 8, <= 256");
  if (n == 0) return SRNN_OK;
  onehot_kernel<<<blocks_for(n * (q / 8), 256), 256, 0, ST(s)>>>(idx, n, q, static_cast<__nv_bfloat16*>(out));
  SRNN_CUDA(cudaGetLastError());
  return SRNN_OK;
}

extern "C" int srnn_weight_prep(const float* v, const float* g, int32_t R, int32_t A, int32_t B, void* o1,
                                const int64_t* s1, void* o2, const int64_t* s2, float* inv_norm, srnn_stream_t s) {
  SRNN_CHECK_ARG(v && R > 0 && A > 0 && B > 0 && o1 && s1, "weight_prep: bad arguments");
  SRNN_CHECK_ARG(!o2 || s2, "weight_prep: second output needs strides");
  Strides3 a{{s1[0], s1[1], s1[2]}}, b{{0, 0, 0}};
  if (o2) b = Strides3{{s2[0], s2[1], s2[2]}};
  weight_prep_kernel<<<R, 256, 0, ST(s)>>>(v, g, A, B, static_cast<__nv_bfloat16*>(o1), a,
                                          static_cast<__nv_bfloat16*>(o2), b, inv_norm);
  SRNN_CUDA(cudaGetLastError());
  return SRNN_OK;
}

extern "C" int srnn_weight_prep_bwd(const float* dw, const int64_t* st, const float* v, const float* g,
                                    const float* inv_norm, int32_t R, int32_t A, int32_t B, float* dv, float* dg,
                                    srnn_stream_t s) {
  SRNN_CHECK_ARG(dw && st && v && dv && R > 0 && A > 0 && B > 0, "weight_prep_bwd: bad arguments");
  SRNN_CHECK_ARG(!g || (inv_norm && dg), "weight_prep_bwd: weight-normed tensors need inv_norm and dg");
  Strides3 a{{st[0], st[1], st[2]}};
  weight_prep_bwd_kernel<<<R, 256, 0, ST(s)>>>(dw, a, v, g, inv_norm, A, B, dv, dg);
  SRNN_CUDA(cudaGetLastError());
  return SRNN_OK;
}

extern "C" int srnn_pad_cast_bf16(const float* in, int64_t rows, int32_t cols, int64_t ld_in, void* out,
                                  int32_t cols_pad, int64_t ld_out, srnn_stream_t s) {
  SRNN_CHECK_ARG(in && out && rows >= 0 && cols > 0 && cols_pad >= cols, "pad_cast_bf16: bad arguments");
  if (rows == 0) return SRNN_OK;
  pad_cast_kernel<<<blocks_for(rows * cols_pad, 256), 256, 0, ST(s)>>>(in, rows, cols, ld_in,
                                                                      static_cast<__nv_bfloat16*>(out), cols_pad, ld_out);
  SRNN_CUDA(cudaGetLastError());
  return SRNN_OK;
}

extern "C" int srnn_split_bf16(const float* in, int64_t rows, int32_t cols, int64_t ld_in, void* hi, void* lo,
                               int32_t cols_pad, int64_t ld_out, srnn_stream_t s) {
  SRNN_CHECK_ARG(in && hi && lo && rows >= 0 && cols > 0 && cols_pad >= cols, "split_bf16: bad arguments");
  if (rows == 0) return SRNN_OK;
  split_bf16_kernel<<<blocks_for(rows * cols_pad, 256), 256, 0, ST(s)>>>(
      in, rows, cols, ld_in, static_cast<__nv_bfloat16*>(hi), static_cast<__nv_bfloat16*>(lo), cols_pad, ld_out);
  SRNN_CUDA(cudaGetLastError());
  return SRNN_OK;
}

extern "C" int srnn_bf16_to_f32(const void* in, int64_t rows, int32_t cols, int64_t ld_in, float* out, int64_t ld_out,
                                int32_t accumulate, srnn_stream_t s) {
  SRNN_CHECK_ARG(in && out && rows >= 0 && cols > 0, "bf16_to_f32: bad arguments");
  if (rows == 0) return SRNN_OK;
  bf16_to_f32_kernel<<<blocks_for(rows * cols, 256), 256, 0, ST(s)>>>(static_cast<const __nv_bfloat16*>(in), rows, cols,
                                                                     ld_in, out, ld_out, accumulate);
  SRNN_CUDA(cudaGetLastError());
  return SRNN_OK;
}

extern "C" int srnn_mixer_input(const float* utt, const float* table, const int32_t* spk, int32_t batch, int32_t frames,
                                int32_t U, int32_t S, void* out, int32_t k_pad, srnn_stream_t s) {
  SRNN_CHECK_ARG(utt && table && spk && out && batch > 0 && frames > 0 && k_pad >= U + S && k_pad % 8 == 0,
                 "mixer_input: bad arguments");
  mixer_input_kernel<<<blocks_for(static_cast<long long>(batch) * frames * k_pad, 256), 256, 0, ST(s)>>>(
      utt, table, spk, batch, frames, U, S, static_cast<__nv_bfloat16*>(out), k_pad);
  SRNN_CUDA(cudaGetLastError());
  return SRNN_OK;
}

extern "C" int srnn_mixer_input_bwd(const void* d_in, const int32_t* spk, int32_t batch, int32_t frames, int32_t S,
                                    int32_t k_pad, float* d_table, srnn_stream_t s) {
  SRNN_CHECK_ARG(d_in && spk && d_table && batch > 0 && frames > 0, "mixer_input_bwd: bad arguments");
  mixer_input_bwd_kernel<<<batch, 32, 0, ST(s)>>>(static_cast<const __nv_bfloat16*>(d_in), spk, frames, S, k_pad,
                                                 d_table);
  SRNN_CUDA(cudaGetLastError());
  return SRNN_OK;
}

extern "C" int srnn_tier_input(const uint8_t* xq, int64_t xq_ld, int32_t x_off, const float* lut, const float* frames,
                               const float* conds, int32_t batch, int32_t T, int32_t fs, int32_t L, int32_t C, void* out,
                               int32_t k_pad, srnn_stream_t s) {
  SRNN_CHECK_ARG(((xq && lut) || frames) && conds && out && batch > 0 && T > 0 && L > 0 && T % L == 0 && k_pad >= fs + C &&
                     k_pad % 8 == 0,
                 "tier_input: bad arguments (T=%d L=%d fs=%d C=%d k_pad=%d)", T, L, fs, C, k_pad);
  tier_input_kernel<<<blocks_for(static_cast<long long>(batch) * T * k_pad, 256), 256, 0, ST(s)>>>(
      xq, xq_ld, x_off, lut, frames, conds, batch, T, fs, L, C, static_cast<__nv_bfloat16*>(out), k_pad);
  SRNN_CUDA(cudaGetLastError());
  return SRNN_OK;
}

extern "C" int srnn_tier_input_bwd(const void* d_in, int32_t batch, int32_t T, int32_t fs, int32_t L, int32_t C,
                                   int32_t k_pad, float* dconds, srnn_stream_t s) {
  SRNN_CHECK_ARG(d_in && dconds && batch > 0 && T > 0 && L > 0 && T % L == 0, "tier_input_bwd: bad arguments");
  tier_input_bwd_kernel<<<blocks_for(static_cast<long long>(batch) * L * C, 256), 256, 0, ST(s)>>>(
      static_cast<const __nv_bfloat16*>(d_in), batch, T, fs, L, C, k_pad, dconds);
  SRNN_CUDA(cudaGetLastError());
  return SRNN_OK;
}

extern "C" int srnn_repeat_rows(const void* in, int64_t rows, int32_t cols, int64_t ld_in, int32_t rep, void* out,
                                int64_t ld_out, srnn_stream_t s) {
  SRNN_CHECK_ARG(in && out && rows > 0 && cols % 8 == 0 && ld_in % 8 == 0 && ld_out % 8 == 0 && rep > 0,
                 "repeat_rows: cols/ld must be multiples of 8");
  repeat_rows_kernel<<<blocks_for(rows * rep * (cols / 8), 256), 256, 0, ST(s)>>>(
      static_cast<const __nv_bfloat16*>(in), rows, cols, ld_in, rep, static_cast<__nv_bfloat16*>(out), ld_out);
  SRNN_CUDA(cudaGetLastError());
  return SRNN_OK;
}

extern "C" int srnn_repeat_rows_bwd(const void* dout, int64_t rows, int32_t cols, int64_t ld_dout, int32_t rep,
                                    void* din, int64_t ld_din, srnn_stream_t s) {
  SRNN_CHECK_ARG(dout && din && rows > 0 && cols % 8 == 0 && ld_dout % 8 == 0 && ld_din % 8 == 0 && rep > 0,
                 "repeat_rows_bwd: cols/ld must be multiples of 8");
  repeat_rows_bwd_kernel<<<blocks_for(rows * (cols / 8), 256), 256, 0, ST(s)>>>(
      static_cast<const __nv_bfloat16*>(dout), rows, cols, ld_dout, rep, static_cast<__nv_bfloat16*>(din), ld_din);
  SRNN_CUDA(cudaGetLastError());
  return SRNN_OK;
}

extern "C" int srnn_colsum(const void* in, int64_t rows, int32_t cols, int64_t ld, float* out, srnn_stream_t s) {
  SRNN_CHECK_ARG(in && out && rows > 0 && cols > 0, "colsum: bad arguments");
  SRNN_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * cols, ST(s)));
  const int col_blocks = (cols + 255) / 256;
  long long slabs = (8LL * sm_count() + col_blocks - 1) / col_blocks;
  if (slabs > (rows + 127) / 128) slabs = (rows + 127) / 128;
  if (slabs < 1) slabs = 1;
  const long long rpb = (rows + slabs - 1) / slabs;
  colsum_kernel<<<dim3(col_blocks, static_cast<unsigned>(slabs)), dim3(32, 8), 0, ST(s)>>>(
      static_cast<const __nv_bfloat16*>(in), rows, cols, ld, rpb, out);
  SRNN_CUDA(cudaGetLastError());
  return SRNN_OK;
}

extern "C" int srnn_state_select(const float* carried, const float* h0, const uint8_t* use_carry, int32_t batch,
                                 int32_t H, float* h_state, void* h_ext, int64_t ext_ld, srnn_stream_t s) {
  SRNN_CHECK_ARG(h0 && use_carry && h_state && batch > 0 && H > 0, "state_select: bad arguments");
  state_select_kernel<<<blocks_for(static_cast<long long>(batch) * H, 256), 256, 0, ST(s)>>>(
      carried, h0, use_carry, batch, H, h_state, static_cast<__nv_bfloat16*>(h_ext), ext_ld);
  SRNN_CUDA(cudaGetLastError());
  return SRNN_OK;
}

extern "C" int srnn_state_select_bwd(const float* dh, const uint8_t* use_carry, int32_t batch, int32_t H, float* d_h0,
                                     srnn_stream_t s) {
  SRNN_CHECK_ARG(dh && use_carry && d_h0 && batch > 0 && H > 0, "state_select_bwd: bad arguments");
  state_select_bwd_kernel<<<blocks_for(H, 128), 128, 0, ST(s)>>>(dh, use_carry, batch, H, d_h0);
  SRNN_CUDA(cudaGetLastError());
  return SRNN_OK;
}

extern "C" int srnn_masked_nll_mean(const float* lp, const uint8_t* valid, int64_t rows, int32_t rows_per_slot,
                                    float* out2, srnn_stream_t s) {
  SRNN_CHECK_ARG(lp && valid && out2 && rows > 0 && rows_per_slot > 0, "masked_nll_mean: bad arguments");
  SRNN_CUDA(cudaMemsetAsync(out2, 0, 2 * sizeof(float), ST(s)));
  unsigned blocks = blocks_for(rows, 256 * 8);
  if (blocks > 4u * sm_count()) blocks = 4u * sm_count();
  masked_sum_kernel<<<blocks, 256, 0, ST(s)>>>(lp, valid, rows, rows_per_slot, out2);
  nll_finalize_kernel<<<1, 1, 0, ST(s)>>>(out2);
  SRNN_CUDA(cudaGetLastError());
  return SRNN_OK;
}

extern "C" int srnn_adam_clipped(float* w, const float* g, float* m, float* v, int64_t n, double lr, double b1,
                                 double b2, double eps, int32_t step, double gscale, srnn_stream_t s) {
  SRNN_CHECK_ARG(w && g && m && v && n >= 0 && step >= 1, "adam_clipped: bad arguments");
  if (n == 0) return SRNN_OK;
  // bias corrections in double, like torch.optim.Adam's Python scalars
  const double bc1 = 1.0 - pow(b1, static_cast<double>(step));
  const double bc2 = sqrt(1.0 - pow(b2, static_cast<double>(step)));
  adam_clipped_kernel<<<blocks_for(n, 256), 256, 0, ST(s)>>>(w, g, m, v, n, static_cast<float>(lr), static_cast<float>(b1),
                                                            static_cast<float>(b2), static_cast<float>(eps),
                                                            static_cast<float>(bc1), static_cast<float>(bc2),
                                                            static_cast<float>(gscale));
  SRNN_CUDA(cudaGetLastError());
  return SRNN_OK;
}
