// Per-sample kernels of batched autoregressive generation (SampleRNNModel.test, model.py:289-351;
// BASELINE config 5).  Both are tiny (a few hundred rows): what matters is that each sample step needs
// ONE launch for the embedding side and ONE for the draw, instead of the dozen small library kernels
// (one-hot, cat, multinomial's checks and reductions, window shift) they replace.
#include <math.h>

#include "common.cuh"

namespace srnn {

// out[b, :] = act( sum_k table[k*q + idx[b, k], :] + pre[b, :] ), 8 columns (16 bytes) per thread.
// model.py:192-200 restated: the embedding + conv1d(k = r0) + the embedding block of comb_layer are linear
// in the one-hot codes, so for fixed weights they collapse into r0 tables of q rows; `pre` carries the
// conditioning and upper-tier blocks of comb_layer (constant over a frame) and its bias.
__global__ void embed_sum_kernel(const __nv_bfloat16* __restrict__ table, const uint8_t* __restrict__ idx,
                                 long long idx_ld, int r0, int q, int hidden, const __nv_bfloat16* __restrict__ pre,
                                 long long pre_ld, int relu, __nv_bfloat16* __restrict__ out, long long out_ld) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y;
  const int c8 = blockIdx.x * blockDim.x + threadIdx.x;
  if (c8 * 8 >= hidden) return;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (pre) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(pre + b * pre_ld) + c8);
    acc[0] = bf16_lo(u.x); acc[1] = bf16_hi(u.x); acc[2] = bf16_lo(u.y); acc[3] = bf16_hi(u.y);
    acc[4] = bf16_lo(u.z); acc[5] = bf16_hi(u.z); acc[6] = bf16_lo(u.w); acc[7] = bf16_hi(u.w);
  }
  for (int k = 0; k < r0; ++k) {
    const int code = idx[b * idx_ld + k];
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(table + (static_cast<long long>(k) * q + code) * hidden) + c8);
    acc[0] += bf16_lo(u.x); acc[1] += bf16_hi(u.x); acc[2] += bf16_lo(u.y); acc[3] += bf16_hi(u.y);
    acc[4] += bf16_lo(u.z); acc[5] += bf16_hi(u.z); acc[6] += bf16_lo(u.w); acc[7] += bf16_hi(u.w);
  }
  if (relu) {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaxf(acc[i], 0.f);
  }
  uint4 o;
  o.x = pack_bf16x2(acc[0], acc[1]);
  o.y = pack_bf16x2(acc[2], acc[3]);
  o.z = pack_bf16x2(acc[4], acc[5]);
  o.w = pack_bf16x2(acc[6], acc[7]);
  reinterpret_cast<uint4*>(out + b * out_ld)[c8] = o;
}

// One warp per utterance.  The row is either log-probabilities or raw logits (`normalise`: the log-softmax of
// model.py:203 is done here, and optionally written to logp_out); then draw from the distribution by inverse CDF
// with the uniform u[b] (model.py:346-348: multinomial of the softmax), or take the arg-max when u is null; append
// the code to the utterance's window of the last `win_len` samples (shift left by one) and store it in out[b].
constexpr int SAMPLE_PER = 8;                          // classes per lane (q <= 256)

// Philox4x32-10 (Salmon et al., SC'11): counter-based, so the draw of utterance b at sample step n is a pure function of
// (seed, n, b) - reproducible whatever the launch geometry, and it lives inside the captured step program.
__device__ __forceinline__ uint32_t philox_uniform_bits(unsigned long long seed, unsigned long long step, uint32_t b) {
  uint32_t c0 = b, c1 = 0u, c2 = static_cast<uint32_t>(step), c3 = static_cast<uint32_t>(step >> 32);
  uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c0;
}

__global__ void sample_kernel(const float* __restrict__ in, long long ld, int batch, int q, int normalise,
                              float* __restrict__ logp_out, long long ld_out, const float* __restrict__ u,
                              unsigned long long* __restrict__ rng, uint8_t* __restrict__ win, int win_len,
                              uint8_t* __restrict__ out, long long out_ld, const __nv_bfloat16* __restrict__ table,
                              int r0, int hidden, const __nv_bfloat16* __restrict__ pre_next, long long pre_ld,
                              __nv_bfloat16* __restrict__ h1_next, long long h1_ld) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  // device-side draws: rng = {seed, step, blocks done}; every block reads the step, the last one to finish advances it
  unsigned long long rng_seed = 0, rng_step = 0;
  if (rng) {
    rng_seed = rng[0];
    rng_step = *reinterpret_cast<volatile unsigned long long*>(rng + 1);
    __syncthreads();                                  // all warps of the block have read the step
    if (threadIdx.x == 0) {
      __threadfence();
      if (atomicAdd(rng + 2, 1ull) == gridDim.x - 1) {
        rng[2] = 0ull;
        rng[1] = rng_step + 1ull;
      }
    }
  }
  if (b >= batch) return;
  const float* row = in + b * ld;
  float x[SAMPLE_PER];                                 // log-probabilities of classes lane*8 .. lane*8+7
#pragma unroll
  for (int i = 0; i < SAMPLE_PER; ++i) {
    const int c = lane * SAMPLE_PER + i;
    x[i] = c < q ? row[c] : -INFINITY;
  }
  if (normalise) {
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < SAMPLE_PER; ++i) mx = fmaxf(mx, x[i]);
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < SAMPLE_PER; ++i) sum += expf(x[i] - mx);
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float lse = mx + logf(sum);
#pragma unroll
    for (int i = 0; i < SAMPLE_PER; ++i) x[i] -= lse;
  }
  if (logp_out) {
#pragma unroll
    for (int i = 0; i < SAMPLE_PER; ++i) {
      const int c = lane * SAMPLE_PER + i;
      if (c < q) logp_out[b * ld_out + c] = x[i];
    }
  }
  int pick;
  if (u || rng) {
    float pr[SAMPLE_PER];
    float local = 0.f;
#pragma unroll
    for (int i = 0; i < SAMPLE_PER; ++i) {
      pr[i] = expf(x[i]);                              // exp(-inf) = 0 for the padding classes
      local += pr[i];
    }
    float incl = local;                                // inclusive scan of the lane sums
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    const float total = __shfl_sync(0xffffffffu, incl, 31);
    // [0,1) with 24 random bits when drawn on the device (same resolution as torch's float uniform)
    const float ub = u ? u[b] : static_cast<float>(philox_uniform_bits(rng_seed, rng_step, static_cast<uint32_t>(b)) >> 8) *
                                    (1.0f / 16777216.0f);
    const float target = ub * total;
    // first lane whose inclusive sum exceeds the target (the last lane with mass if rounding leaves none)
    const unsigned ahead = __ballot_sync(0xffffffffu, incl > target);
    const unsigned mass = __ballot_sync(0xffffffffu, local > 0.f);
    const int src = ahead ? __ffs(ahead) - 1 : (mass ? 31 - __clz(mass) : 0);
    int cand = lane * SAMPLE_PER;
    {
      float run = incl - local;
      int last_nz = -1;
      bool found = false;
#pragma unroll
      for (int i = 0; i < SAMPLE_PER; ++i) {
        if (pr[i] > 0.f) last_nz = lane * SAMPLE_PER + i;
        run += pr[i];
        if (!found && run > target) {
          cand = lane * SAMPLE_PER + i;
          found = true;
        }
      }
      if (!found && last_nz >= 0) cand = last_nz;
    }
    pick = __shfl_sync(0xffffffffu, cand, src);
  } else {
    float best = -INFINITY;
    int bi = lane * SAMPLE_PER;
#pragma unroll
    for (int i = 0; i < SAMPLE_PER; ++i) {
      if (x[i] > best) {
        best = x[i];
        bi = lane * SAMPLE_PER + i;
      }
    }
    for (int o = 16; o; o >>= 1) {                      // ties resolve to the lowest class index
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) {
        best = ob;
        bi = oi;
      }
    }
    pick = bi;
  }
  if (win) {
    uint8_t* w = win + static_cast<long long>(b) * win_len;
    for (int base = 0; base < win_len; base += 32) {
      const int i = base + lane;
      uint8_t v = 0;
      if (i < win_len) v = (i + 1 < win_len) ? w[i + 1] : static_cast<uint8_t>(pick);
      __syncwarp();
      if (i < win_len) w[i] = v;
      __syncwarp();
    }
  }
  if (out && lane == 0) out[b * out_ld] = static_cast<uint8_t>(pick);
  // Fused head of the NEXT sample step (when its frame-constant term is already known): the embedding side of
  // comb_layer for the window that now ends with the code just drawn - srnn_embed_sum without a launch of its own.
  if (h1_next) {
    __syncwarp();
    const uint8_t* w = win + static_cast<long long>(b) * win_len + (win_len - r0);
    for (int c8 = lane; c8 * 8 < hidden; c8 += 32) {
      float acc[8];
      {
        const uint4 u4 = __ldg(reinterpret_cast<const uint4*>(pre_next + b * pre_ld) + c8);
        acc[0] = bf16_lo(u4.x); acc[1] = bf16_hi(u4.x); acc[2] = bf16_lo(u4.y); acc[3] = bf16_hi(u4.y);
        acc[4] = bf16_lo(u4.z); acc[5] = bf16_hi(u4.z); acc[6] = bf16_lo(u4.w); acc[7] = bf16_hi(u4.w);
      }
      for (int k = 0; k < r0; ++k) {
        const int code = (k == r0 - 1) ? pick : static_cast<int>(w[k]);
        const uint4 u4 = __ldg(reinterpret_cast<const uint4*>(table + (static_cast<long long>(k) * q + code) * hidden) + c8);
        acc[0] += bf16_lo(u4.x); acc[1] += bf16_hi(u4.x); acc[2] += bf16_lo(u4.y); acc[3] += bf16_hi(u4.y);
        acc[4] += bf16_lo(u4.z); acc[5] += bf16_hi(u4.z); acc[6] += bf16_lo(u4.w); acc[7] += bf16_hi(u4.w);
      }
      uint4 o;
      o.x = pack_bf16x2(fmaxf(acc[0], 0.f), fmaxf(acc[1], 0.f));
      o.y = pack_bf16x2(fmaxf(acc[2], 0.f), fmaxf(acc[3], 0.f));
      o.z = pack_bf16x2(fmaxf(acc[4], 0.f), fmaxf(acc[5], 0.f));
      o.w = pack_bf16x2(fmaxf(acc[6], 0.f), fmaxf(acc[7], 0.f));
      reinterpret_cast<uint4*>(h1_next + b * h1_ld)[c8] = o;
    }
  }
}

// <<<>>> with the programmatic-stream-serialization attribute when srnn_set_pdl(1) is in effect
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace srnn

using namespace srnn;

extern "C" int srnn_embed_sum(const void* table, const uint8_t* idx, int64_t idx_ld, int32_t batch, int32_t r0,
                              int32_t q, int32_t hidden, const void* pre, int64_t pre_ld, int32_t relu, void* out,
                              int64_t out_ld, srnn_stream_t s) {
  SRNN_CHECK_ARG(table && idx && out && batch > 0 && r0 > 0 && q > 0 && q <= 256, "embed_sum: bad arguments");
  SRNN_CHECK_ARG(hidden % 8 == 0 && out_ld % 8 == 0 && (!pre || pre_ld % 8 == 0),
                 "embed_sum: hidden and the row strides must be multiples of 8");
  SRNN_CHECK_ARG(((reinterpret_cast<uintptr_t>(table) | reinterpret_cast<uintptr_t>(out) |
                   reinterpret_cast<uintptr_t>(pre)) & 15) == 0, "embed_sum: pointers must be 16-byte aligned");
  const int threads = 128;
  dim3 grid((hidden / 8 + threads - 1) / threads, batch);
  SRNN_CUDA(launch_pdl(embed_sum_kernel, grid, dim3(threads), static_cast<cudaStream_t>(s),
                       static_cast<const __nv_bfloat16*>(table), idx, static_cast<long long>(idx_ld), r0, q, hidden,
                       static_cast<const __nv_bfloat16*>(pre), static_cast<long long>(pre_ld), relu,
                       static_cast<__nv_bfloat16*>(out), static_cast<long long>(out_ld)));
  return SRNN_OK;
}

static int sample_launch(const float* in, int64_t ld, int32_t batch, int32_t q, int32_t normalise, float* logp_out,
                         int64_t ld_out, const float* u, uint64_t* rng_state, uint8_t* win, int32_t win_len, uint8_t* out,
                         int64_t out_ld, const void* table, int32_t r0, int32_t hidden, const void* pre_next, int64_t pre_ld,
                         void* h1_next, int64_t h1_ld, srnn_stream_t s) {
  SRNN_CHECK_ARG(in && batch > 0 && q > 0 && q <= 32 * SAMPLE_PER && (win || out || logp_out),
                 "sample_categorical: bad arguments");
  SRNN_CHECK_ARG(!win || win_len > 0, "sample_categorical: win_len must be positive");
  const int warps = 4;
  SRNN_CUDA(launch_pdl(sample_kernel, dim3((batch + warps - 1) / warps), dim3(warps * 32), static_cast<cudaStream_t>(s),
                       in, static_cast<long long>(ld), batch, q, normalise, logp_out, static_cast<long long>(ld_out), u,
                       reinterpret_cast<unsigned long long*>(rng_state), win, win_len, out, static_cast<long long>(out_ld),
                       static_cast<const __nv_bfloat16*>(table), r0, hidden, static_cast<const __nv_bfloat16*>(pre_next),
                       static_cast<long long>(pre_ld), static_cast<__nv_bfloat16*>(h1_next), static_cast<long long>(h1_ld)));
  return SRNN_OK;
}

extern "C" int srnn_sample_categorical(const float* in, int64_t ld, int32_t batch, int32_t q, int32_t normalise,
                                       float* logp_out, int64_t ld_out, const float* u, uint64_t* rng_state, uint8_t* win,
                                       int32_t win_len, uint8_t* out, int64_t out_ld, srnn_stream_t s) {
  return sample_launch(in, ld, batch, q, normalise, logp_out, ld_out, u, rng_state, win, win_len, out, out_ld, nullptr, 0, 0,
                       nullptr, 0, nullptr, 0, s);
}

extern "C" int srnn_sample_embed(const float* in, int64_t ld, int32_t batch, int32_t q, int32_t normalise, float* logp_out,
                                 int64_t ld_out, const float* u, uint64_t* rng_state, uint8_t* win, int32_t win_len,
                                 uint8_t* out, int64_t out_ld, const void* table, int32_t r0, int32_t hidden,
                                 const void* pre_next, int64_t pre_ld, void* h1_next, int64_t h1_ld, srnn_stream_t s) {
  SRNN_CHECK_ARG(table && pre_next && h1_next && win && r0 > 0 && r0 <= win_len && hidden > 0 && hidden % 8 == 0 &&
                 pre_ld % 8 == 0 && h1_ld % 8 == 0, "sample_embed: bad arguments");
  SRNN_CHECK_ARG(((reinterpret_cast<uintptr_t>(table) | reinterpret_cast<uintptr_t>(pre_next) |
                   reinterpret_cast<uintptr_t>(h1_next)) & 15) == 0, "sample_embed: pointers must be 16-byte aligned");
  return sample_launch(in, ld, batch, q, normalise, logp_out, ld_out, u, rng_state, win, win_len, out, out_ld, table, r0, hidden,
                       pre_next, pre_ld, h1_next, h1_ld, s);
}
