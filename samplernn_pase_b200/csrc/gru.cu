// Persistent recurrent GRU kernels for sm_100a (replace torch.nn.GRU, model.py:110,152, and its
// autograd backward).
//
// One cooperative launch runs all T timesteps of a tier.  The hidden units are partitioned over
// G = H/U CTAs; each CTA keeps its slice of W_hh (forward: the 3U gate rows of its U units;
// backward: its U rows of W_hh^T) resident in shared memory for the whole launch, in the
// 128-byte-swizzled K-major layout tcgen05 consumes.  Per timestep every CTA
//   1. waits on a grid-wide arrival counter (release/acquire through L2) for h_{t-1} (resp. the
//      gate gradients of step t+1) of all CTAs,
//   2. streams that [batch, K] bf16 matrix through a TMA ring and multiplies it against the
//      resident weight slice with tcgen05.mma (M = 64 or 128 batch rows, fp32 accumulate in TMEM),
//   3. finishes the gate math in the epilogue warps (fp32), writes its U columns of h_t (bf16
//      exchange copy + saved gates) and arrives on the counter.
// The fp32 recurrent state of a unit never leaves the registers of its owner thread.
#include "common.cuh"

namespace srnn {

constexpr int GRU_THREADS = 192;   // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue

struct GruParams {
  int batch, steps, hidden, kblocks, ctas;
  const __nv_bfloat16* gi;
  const float* b_hh;
  __nv_bfloat16* h_ext;
  float* h_state;
  __nv_bfloat16* gates;
  const __nv_bfloat16* dh_out;
  __nv_bfloat16* dgi;
  __nv_bfloat16* dgh;
  float* dh0;
  uint32_t* sync;
};

template <int MT>
struct GruCfg {
  static constexpr int U = MT == 64 ? 8 : 16;          // hidden units per CTA
  static constexpr int RING = MT == 64 ? 8 : 4;         // TMA ring stages
  static constexpr int SLOT_BYTES = MT * 128;           // [MT rows][64 bf16]
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

template <int U>
__device__ __forceinline__ void load_bf16_vec(const __nv_bfloat16* p, float (&out)[U]) {
#pragma unroll
  for (int i = 0; i < U / 8; ++i) {
    const uint4 u = *reinterpret_cast<const uint4*>(p + i * 8);
    out[i * 8 + 0] = bf16_lo(u.x); out[i * 8 + 1] = bf16_hi(u.x);
    out[i * 8 + 2] = bf16_lo(u.y); out[i * 8 + 3] = bf16_hi(u.y);
    out[i * 8 + 4] = bf16_lo(u.z); out[i * 8 + 5] = bf16_hi(u.z);
    out[i * 8 + 6] = bf16_lo(u.w); out[i * 8 + 7] = bf16_hi(u.w);
  }
}
template <int U>
__device__ __forceinline__ void store_bf16_vec(__nv_bfloat16* p, const float (&v)[U]) {
#pragma unroll
  for (int i = 0; i < U / 8; ++i) {
    uint4 u;
    u.x = pack_bf16x2(v[i * 8 + 0], v[i * 8 + 1]);
    u.y = pack_bf16x2(v[i * 8 + 2], v[i * 8 + 3]);
    u.z = pack_bf16x2(v[i * 8 + 4], v[i * 8 + 5]);
    u.w = pack_bf16x2(v[i * 8 + 6], v[i * 8 + 7]);
    *reinterpret_cast<uint4*>(p + i * 8) = u;
  }
}
template <int U>
__device__ __forceinline__ void tmem_ld_units(uint32_t taddr, float (&out)[U]) {
  if constexpr (U == 8) {
    uint32_t v[8];
    tmem_ld8(taddr, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i] = __uint_as_float(v[i]);
  } else {
    uint32_t v[16];
    tmem_ld16(taddr, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) out[i] = __uint_as_float(v[i]);
  }
}

__device__ __forceinline__ void grid_wait(const uint32_t* counter, uint32_t target) {
  uint32_t spins = 0;
  while (ld_acquire_gpu(counter) < target) {
    if (++spins > (1u << 24)) __trap();
  }
}

// NG = number of gate row-groups resident per CTA (3 forward: r,z,n rows of W_hh; 1 backward: rows of W_hh^T)
template <int MT, bool BWD>
__global__ void __launch_bounds__(GRU_THREADS, 1)
gru_kernel(const __grid_constant__ CUtensorMap tma_w, const __grid_constant__ CUtensorMap tma_x, const GruParams p) {
  using Cfg = GruCfg<MT>;
  constexpr int U = Cfg::U;
  constexpr int NG = BWD ? 1 : 3;
  constexpr int NCOLS = NG * U;                      // MMA N
  constexpr uint32_t TMEM_COLS = NCOLS <= 32 ? 32 : 64;
  constexpr uint32_t IDESC = idesc_bf16(MT, NCOLS, false, false);
  constexpr int RING = Cfg::RING;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int KB = p.kblocks;                          // K blocks of 64 (K = H forward, 3H backward)
  const int wblock = NCOLS * 128;                    // bytes of the resident weight per K block
  uint8_t* sw = smem;
  uint8_t* ring = smem + ((KB * wblock + 1023) & ~1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + RING * Cfg::SLOT_BYTES);
  uint64_t* wfull = bars;
  uint64_t* full = bars + 1;
  uint64_t* empty = full + RING;
  uint64_t* acc_full = empty + RING;
  uint64_t* acc_empty = acc_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int H = p.hidden, T = p.steps, B = p.batch;
  const int u0 = blockIdx.x * U;
  const uint32_t G = gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_w);
    tma_prefetch_desc(&tma_x);
    mbar_init(wfull, 1);
    for (int s = 0; s < RING; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 4);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // number of matmul rounds: forward T (round s consumes h_{s-1} = slot s); backward T (rounds 1..T
  // consume the gate gradients of step T-s; round 0 has no recurrent input)
  if (warp == 0) {
    if (lane == 0) {
      // resident weights
      mbar_expect_tx(wfull, static_cast<uint32_t>(KB * wblock));
      for (int kb = 0; kb < KB; ++kb)
        for (int g = 0; g < NG; ++g)
          tma_load_2d(sw + kb * wblock + g * U * 128, &tma_w, wfull, kb * 64, (BWD ? 0 : g * H) + u0);
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t slot_bytes = static_cast<uint32_t>(B) * 128u;
      const int rounds = BWD ? T + 1 : T;
      for (int s = 0; s < rounds; ++s) {
        if (BWD && s == 0) continue;
        if (s > 0) {
          grid_wait(p.sync, G * static_cast<uint32_t>(s));
          fence_proxy_async_all();
        }
        const int slot = BWD ? (T - s) : s;        // time slot of the exchange buffer to read
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], slot_bytes);
          tma_load_3d(ring + stage * Cfg::SLOT_BYTES, &tma_x, &full[stage], kb * 64, slot, 0);
          if (++stage == RING) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      mbar_wait(wfull, 0);
      int stage = 0;
      uint32_t phase = 0, acc_phase = 0;
      const int rounds = BWD ? T + 1 : T;
      for (int s = 0; s < rounds; ++s) {
        if (BWD && s == 0) continue;
        mbar_wait(acc_empty, acc_phase ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(ring + stage * Cfg::SLOT_BYTES);
          const uint32_t b_addr = smem_u32(sw + kb * wblock);
#pragma unroll
          for (int k16 = 0; k16 < 4; ++k16) {
            umma_bf16(tmem_base, smem_desc_sw128(a_addr + k16 * 32, 16, 1024),
                      smem_desc_sw128(b_addr + k16 * 32, 16, 1024), IDESC, (kb > 0 || k16 > 0) ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
          if (++stage == RING) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(acc_full);
        acc_phase ^= 1;
      }
    }
    __syncwarp();
  } else {
    // ------------------------------ epilogue: gate math ------------------------------
    const int q = warp & 3;                           // TMEM lane quadrant of this warp
    int row;
    bool row_ok;
    if (MT == 128) {
      row = q * 32 + lane;
      row_ok = row < B;
    } else {
      row = q * 16 + lane;                            // M=64: rows 16q..16q+15 live in lanes 32q..32q+15
      row_ok = lane < 16 && row < B;
    }
    const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint32_t acc_phase = 0;

    if (!BWD) {
      float h[U], bhr[U], bhz[U], bhn[U];
#pragma unroll
      for (int i = 0; i < U; ++i) {
        h[i] = row_ok ? p.h_state[static_cast<long long>(row) * H + u0 + i] : 0.f;
        bhr[i] = p.b_hh[u0 + i];
        bhz[i] = p.b_hh[H + u0 + i];
        bhn[i] = p.b_hh[2 * H + u0 + i];
      }
      for (int t = 0; t < T; ++t) {
        const long long rt = static_cast<long long>(row) * T + t;
        float gr[U], gz[U], gn[U];
        if (row_ok) {
          const __nv_bfloat16* gp = p.gi + rt * 3 * H + u0;
          load_bf16_vec<U>(gp, gr);
          load_bf16_vec<U>(gp + H, gz);
          load_bf16_vec<U>(gp + 2 * H, gn);
        }
        mbar_wait(acc_full, acc_phase);
        acc_phase ^= 1;
        tc_fence_after();
        float dr[U], dz[U], dn[U];
        tmem_ld_units<U>(t_addr, dr);
        tmem_ld_units<U>(t_addr + U, dz);
        tmem_ld_units<U>(t_addr + 2 * U, dn);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty);
        if (row_ok) {
          float r[U], z[U], n[U], hn[U];
#pragma unroll
          for (int i = 0; i < U; ++i) {
            r[i] = sigmoidf_(gr[i] + dr[i] + bhr[i]);
            z[i] = sigmoidf_(gz[i] + dz[i] + bhz[i]);
            hn[i] = dn[i] + bhn[i];
            n[i] = tanhf(gn[i] + r[i] * hn[i]);
            h[i] = (1.f - z[i]) * n[i] + z[i] * h[i];
          }
          store_bf16_vec<U>(p.h_ext + (static_cast<long long>(row) * (T + 1) + t + 1) * H + u0, h);
          if (p.gates) {
            __nv_bfloat16* sp = p.gates + rt * 4 * H + u0;
            store_bf16_vec<U>(sp, r);
            store_bf16_vec<U>(sp + H, z);
            store_bf16_vec<U>(sp + 2 * H, n);
            store_bf16_vec<U>(sp + 3 * H, hn);
          }
        }
        // publish h_t: all epilogue threads' stores -> one release arrival per CTA
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (warp == 2 && lane == 0) {
          __threadfence();
          red_release_gpu_add(p.sync, 1u);
        }
      }
      if (row_ok) {
#pragma unroll
        for (int i = 0; i < U; ++i) p.h_state[static_cast<long long>(row) * H + u0 + i] = h[i];
      }
    } else {
      float carry[U];
#pragma unroll
      for (int i = 0; i < U; ++i) carry[i] = 0.f;
      for (int s = 0; s <= T; ++s) {
        const int t = T - 1 - s;
        const long long rt = static_cast<long long>(row) * T + t;
        float dh[U], r[U], z[U], n[U], hn[U], hp[U];
        if (row_ok && t >= 0) {
          load_bf16_vec<U>(p.dh_out + rt * H + u0, dh);
          const __nv_bfloat16* sp = p.gates + rt * 4 * H + u0;
          load_bf16_vec<U>(sp, r);
          load_bf16_vec<U>(sp + H, z);
          load_bf16_vec<U>(sp + 2 * H, n);
          load_bf16_vec<U>(sp + 3 * H, hn);
          load_bf16_vec<U>(p.h_ext + (static_cast<long long>(row) * (T + 1) + t) * H + u0, hp);
        }
        float d[U];
#pragma unroll
        for (int i = 0; i < U; ++i) d[i] = 0.f;
        if (s > 0) {
          mbar_wait(acc_full, acc_phase);
          acc_phase ^= 1;
          tc_fence_after();
          tmem_ld_units<U>(t_addr, d);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty);
        }
        if (t < 0) {
          if (row_ok) {
#pragma unroll
            for (int i = 0; i < U; ++i) p.dh0[static_cast<long long>(row) * H + u0 + i] = carry[i] + d[i];
          }
          break;
        }
        if (row_ok) {
          float gr[U], gz[U], gn[U], ghn[U];
#pragma unroll
          for (int i = 0; i < U; ++i) {
            const float dht = dh[i] + carry[i] + d[i];
            const float dn_ = dht * (1.f - z[i]);
            const float dz_ = dht * (hp[i] - n[i]);
            carry[i] = dht * z[i];
            const float dn_pre = dn_ * (1.f - n[i] * n[i]);
            const float dr_ = dn_pre * hn[i];
            gr[i] = dr_ * r[i] * (1.f - r[i]);
            gz[i] = dz_ * z[i] * (1.f - z[i]);
            gn[i] = dn_pre;
            ghn[i] = dn_pre * r[i];
          }
          __nv_bfloat16* gip = p.dgi + rt * 3 * H + u0;
          store_bf16_vec<U>(gip, gr);
          store_bf16_vec<U>(gip + H, gz);
          store_bf16_vec<U>(gip + 2 * H, gn);
          __nv_bfloat16* ghp = p.dgh + rt * 3 * H + u0;
          store_bf16_vec<U>(ghp, gr);
          store_bf16_vec<U>(ghp + H, gz);
          store_bf16_vec<U>(ghp + 2 * H, ghn);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (warp == 2 && lane == 0) {
          __threadfence();
          red_release_gpu_add(p.sync, 1u);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int MT, bool BWD>
static int launch_gru(const srnn_gru_args* a, cudaStream_t stream) {
  using Cfg = GruCfg<MT>;
  constexpr int U = Cfg::U;
  const int H = a->hidden, T = a->steps, B = a->batch;
  SRNN_CHECK_ARG(H % U == 0, "gru: hidden (%d) must be a multiple of %d for batch tile %d", H, U, MT);
  const int ctas = H / U;
  SRNN_CHECK_ARG(ctas <= sm_count(), "gru: hidden/%d = %d CTAs exceeds the SM count %d", U, ctas, sm_count());
  const int K = BWD ? 3 * H : H;
  const int KB = (K + 63) / 64;
  const int ncols = (BWD ? 1 : 3) * U;
  const size_t wbytes = ((size_t)KB * ncols * 128 + 1023) & ~(size_t)1023;
  const size_t smem = wbytes + (size_t)Cfg::RING * Cfg::SLOT_BYTES + 512 + 1024;
  SRNN_CHECK_ARG(smem <= 227 * 1024, "gru: resident weight slice does not fit shared memory (%zu bytes)", smem);

  CUtensorMap tw, tx;
  {
    // forward: W_hh [3H, H] rows = gate rows; backward: W_hh^T [H, 3H] rows = units
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)(BWD ? H : 3 * H)};
    const uint64_t strides[1] = {(uint64_t)K * 2};
    const uint32_t box[2] = {64, (uint32_t)U};
    int rc = make_tmap_bf16(&tw, a->w_hh, 2, dims, strides, box, true);
    if (rc) return rc;
  }
  {
    // exchange buffer: forward h_ext [B, T+1, H]; backward dgh [B, T, 3H]
    const uint64_t slots = BWD ? (uint64_t)T : (uint64_t)T + 1;
    const uint64_t dims[3] = {(uint64_t)K, slots, (uint64_t)B};
    const uint64_t strides[2] = {(uint64_t)K * 2, (uint64_t)K * 2 * slots};
    const uint32_t box[3] = {64, 1, (uint32_t)B};
    int rc = make_tmap_bf16(&tx, BWD ? (const void*)a->dgh : (const void*)a->h_ext, 3, dims, strides, box, true);
    if (rc) return rc;
  }
  GruParams p{};
  p.batch = B; p.steps = T; p.hidden = H; p.kblocks = KB; p.ctas = ctas;
  p.gi = static_cast<const __nv_bfloat16*>(a->gi);
  p.b_hh = a->b_hh;
  p.h_ext = static_cast<__nv_bfloat16*>(a->h_ext);
  p.h_state = a->h_state;
  p.gates = static_cast<__nv_bfloat16*>(a->gates);
  p.dh_out = static_cast<const __nv_bfloat16*>(a->dh_out);
  p.dgi = static_cast<__nv_bfloat16*>(a->dgi);
  p.dgh = static_cast<__nv_bfloat16*>(a->dgh);
  p.dh0 = a->dh0;
  p.sync = a->sync;

  auto kern = gru_kernel<MT, BWD>;
  SRNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* args[] = {(void*)&tw, (void*)&tx, (void*)&p};
  SRNN_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3(ctas), dim3(GRU_THREADS), args, smem, stream));
  return SRNN_OK;
}

}  // namespace srnn

using namespace srnn;

static int check_common(const srnn_gru_args* a) {
  SRNN_CHECK_ARG(a != nullptr, "gru: null args");
  SRNN_CHECK_ARG(a->batch > 0 && a->batch <= 128, "gru: batch must be in 1..128 (got %d); split larger batches", a->batch);
  SRNN_CHECK_ARG(a->steps > 0 && a->hidden > 0 && a->hidden % 8 == 0, "gru: bad steps/hidden (%d, %d)", a->steps,
                 a->hidden);
  SRNN_CHECK_ARG(a->w_hh && a->h_ext && a->gates && a->sync, "gru: null buffer");
  return SRNN_OK;
}

extern "C" int srnn_gru_forward(const srnn_gru_args* a, srnn_stream_t stream) {
  int rc = check_common(a);
  if (rc) return rc;
  SRNN_CHECK_ARG(a->gi && a->b_hh && a->h_state, "gru_forward: null buffer");
  if (a->batch <= 64) return launch_gru<64, false>(a, static_cast<cudaStream_t>(stream));
  return launch_gru<128, false>(a, static_cast<cudaStream_t>(stream));
}

extern "C" int srnn_gru_backward(const srnn_gru_args* a, srnn_stream_t stream) {
  int rc = check_common(a);
  if (rc) return rc;
  SRNN_CHECK_ARG(a->dh_out && a->dgi && a->dgh && a->dh0, "gru_backward: null buffer");
  if (a->batch <= 64) return launch_gru<64, true>(a, static_cast<cudaStream_t>(stream));
  return launch_gru<128, true>(a, static_cast<cudaStream_t>(stream));
}
