// tcgen05 / TMEM / TMA GEMM for sm_100a.
//
// One persistent, warp-specialised kernel template serves every dense contraction of the
// training step (reference: cuDNN conv1d/conv_transpose1d and cuBLAS calls behind
// model.py:48,108-109,112,146-147,153-155,168-172,193-202 and their autograd backward):
//
//   op NT : C_i[m,n] = epi(A_i[m,k] . B[n,k]^T)      A, B K-major (activations x weights)
//   op TN : C[m,n]  += sum_i A_i[k,m]^T . B_i[k,n]    A, B MN-major (weight gradients), split-K,
//                                                      fp32 atomics
//   NLL   : NT with N = 256 whose epilogue does log-softmax + NLL (or its backward) straight
//           out of TMEM, so logits never reach HBM (model.py:202-203 + runner.py:52).
//
// Tile 128 x BN x 64 (BN = 128 or 256), bf16 operands staged by TMA with the 128-byte swizzle,
// fp32 accumulators double-buffered in TMEM (2 x BN columns), a ring of 4 (BN=256) or 6 (BN=128)
// smem stages.  Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one thread), warp 2 =
// TMEM allocator, warps 4-7 = epilogue (one TMEM lane quadrant each).
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"

namespace srnn {

constexpr int BM = 128;
constexpr int BK = 64;
// warps 0-3: TMA / MMA / TMEM alloc / spare; then 4 (NLL epilogue: one thread owns a whole row) or 8
// epilogue warps (two per TMEM lane quadrant, each taking half of the tile's columns)
template <int EPI>
__host__ __device__ constexpr int gemm_threads() { return EPI == 1 ? 256 : 384; }

struct GemmParams {
  int m, n, k, batch;
  int a_row_offset, b_row_offset;
  void* c;
  long long ldc, c_batch_stride;
  int c_dtype, n_fold;
  const float* bias;
  const __nv_bfloat16* aux;
  long long ldaux, aux_batch_stride;
  int aux_mode, relu, aux_row_div, max_ctas;
  int c_tma;                       // bf16 C tiles leave through shared memory + TMA stores (tma_c is valid)
  int kb_a1;                       // NT: K blocks taken from the first A operand (the rest from tma_a2); total if single
  int tn_4d;                       // TN: operands described as [batch][MN/64][k][64] - one TMA box per operand and K block
  float* colsum;                   // NT: column sums of the fp32 epilogue result, accumulated with atomics (nullable)
  uint32_t* relu_mask;             // NT + relu (nullable): bit (row, col) = result > 0, one word per row and 32 columns
  const uint32_t* gate_mask;       // NT (nullable): C = acc where the bit is set, else 0 (ReLU gradient gate)
  long long ldmask;                // words per row of either mask
  // schedule
  int tiles_m, tiles_n, kb_per_batch, splits, kb_per_split, total_kb, total_work;
  // NLL epilogue
  int nll_mode;
  const uint8_t* target;
  float* lse;
  float* logp_t;
  float* logp;
  long long ldlogp;
  const float* row_grad;
  const float* g;
  long long ldg;
};

template <int BN>
struct SmemCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int STAGES = BN == 256 ? 4 : 6;
  static constexpr int STG_OFF = STAGES * STAGE;          // store staging: 8 epilogue warps x 2 x (32 rows x 64 B)
  static constexpr int BAR_OFF = STG_OFF + 8 * 2 * 2048;
  static constexpr int BIAS_OFF = BAR_OFF + 256;
  static constexpr int TOTAL = BIAS_OFF + 1024 + 1024;  // + bias tile + alignment slack
};

struct Work {
  int mt, nt, b, kb_begin, kb_end;
};

template <bool TN>
__device__ __forceinline__ Work decode_work(const GemmParams& p, int w) {
  Work wk;
  if (!TN) {
    wk.nt = w % p.tiles_n;
    int t = w / p.tiles_n;
    wk.b = t / p.tiles_m;
    wk.mt = t % p.tiles_m;
    wk.kb_begin = 0;
    wk.kb_end = p.kb_per_batch;
  } else {
    // split-major order: the CTAs that run at the same time work on the SAME K slab of different output
    // tiles, so the A/B rows of that slab are fetched from HBM once and shared through L2 (tile-major order
    // made every CTA stream private rows: 3x the algorithmic DRAM reads in ncu)
    const int tiles = p.tiles_m * p.tiles_n;
    int split = w / tiles;
    int tile = w - split * tiles;
    wk.nt = tile % p.tiles_n;
    wk.mt = tile / p.tiles_n;
    wk.b = 0;
    wk.kb_begin = split * p.kb_per_split;
    wk.kb_end = min(wk.kb_begin + p.kb_per_split, p.total_kb);
  }
  return wk;
}

// EPI: 0 = bias / aux / relu / store, 1 = log-softmax + NLL family, 2 = fp32 atomic accumulate,
//      3 = as 0 plus the column sums of the result (fused bias gradient; its own instantiation so that the plain
//          store epilogue carries neither its registers nor its code)
template <int BN, bool TN, int EPI>
__global__ void __launch_bounds__(gemm_threads<EPI>(), 1)
gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
            const __grid_constant__ CUtensorMap tma_c, const __grid_constant__ CUtensorMap tma_a2, const GemmParams p) {
  using Cfg = SmemCfg<BN>;
  constexpr int S = Cfg::STAGES;
  constexpr uint32_t TMEM_COLS = 2 * BN;
  constexpr uint32_t IDESC = idesc_bf16(BM, BN, TN, TN);
  constexpr int THREADS = gemm_threads<EPI>();
  constexpr bool STORE = EPI == 0 || EPI == 3;
  constexpr bool CSUM = EPI == 3;
  constexpr int EPI_WARPS = THREADS / 32 - 4;
  constexpr int CHUNKS_PER_WARP = (BN / 32) / (EPI_WARPS / 4);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
  uint64_t* empty = full + S;
  uint64_t* tfull = empty + S;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* sbias = reinterpret_cast<float*>(smem + Cfg::BIAS_OFF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    if (STORE && p.c_tma) tma_prefetch_desc(&tma_c);
    if (!TN && p.kb_a1 < p.kb_per_batch) tma_prefetch_desc(&tma_a2);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  if constexpr (EPI == 1) {
    for (int i = threadIdx.x; i < 256; i += THREADS) sbias[i] = p.bias ? p.bias[i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------ TMA producer ------------------------------
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < p.total_work; w += gridDim.x) {
        const Work wk = decode_work<TN>(p, w);
        for (int kb = wk.kb_begin; kb < wk.kb_end; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE;
          uint8_t* sb = sa + Cfg::A_BYTES;
          mbar_expect_tx(&full[stage], Cfg::STAGE);
          if (!TN) {
            if (kb < p.kb_a1) {
              tma_load_3d(sa, &tma_a, &full[stage], kb * BK, wk.mt * BM, wk.b);
            } else {                                   // second A operand: columns [k1, k) of the concatenation
              tma_load_3d(sa, &tma_a2, &full[stage], (kb - p.kb_a1) * BK, wk.mt * BM, wk.b);
            }
            tma_load_2d(sb, &tma_b, &full[stage], kb * BK, wk.nt * BN);
          } else {
            const int bi = kb / p.kb_per_batch;
            const int r0 = (kb - bi * p.kb_per_batch) * BK;
            if (p.tn_4d) {                             // 2 TMA issues per K block instead of 6 (the producer thread
              tma_load_4d(sa, &tma_a, &full[stage], 0, r0 + p.a_row_offset, wk.mt * (BM / 64), bi);   // was the limit)
              tma_load_4d(sb, &tma_b, &full[stage], 0, r0 + p.b_row_offset, wk.nt * (BN / 64), bi);
            } else {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j)
                tma_load_3d(sa + j * 8192, &tma_a, &full[stage], wk.mt * BM + j * 64, r0 + p.a_row_offset, bi);
#pragma unroll
              for (int j = 0; j < BN / 64; ++j)
                tma_load_3d(sb + j * 8192, &tma_b, &full[stage], wk.nt * BN + j * 64, r0 + p.b_row_offset, bi);
            }
          }
          if (++stage == S) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------ MMA issuer ------------------------------
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int w = blockIdx.x; w < p.total_work; w += gridDim.x) {
        const Work wk = decode_work<TN>(p, w);
        if (wk.kb_end <= wk.kb_begin) continue;
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = wk.kb_begin; kb < wk.kb_end; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * Cfg::STAGE);
          const uint32_t b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
          for (int k16 = 0; k16 < BK / 16; ++k16) {
            uint64_t ad, bd;
            if (!TN) {
              ad = smem_desc_sw128(a_addr + k16 * 32, 16, 1024);
              bd = smem_desc_sw128(b_addr + k16 * 32, 16, 1024);
            } else {
              ad = smem_desc_sw128(a_addr + k16 * 2048, 8192, 1024);
              bd = smem_desc_sw128(b_addr + k16 * 2048, 8192, 1024);
            }
            umma_bf16(d_tmem, ad, bd, IDESC, (kb > wk.kb_begin || k16 > 0) ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
          if (++stage == S) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ------------------------------ epilogue ------------------------------
    const int q = warp & 3;
    const int c_begin = ((warp - 4) >> 2) * CHUNKS_PER_WARP;   // this warp's share of the tile's columns
    const int c_end = c_begin + CHUNKS_PER_WARP;
    int acc = 0;
    uint32_t acc_phase = 0;
    int stg_buf = 0;
    // fused bias gradient (p.colsum): lane i of this warp keeps the running sum of column i of each of its chunks over
    // the CTA's tiles that share a column tile; flushed with one coalesced atomic per chunk when the column tile changes
    float csum[CHUNKS_PER_WARP];
    int cs_nt = -1;
#pragma unroll
    for (int i = 0; i < CHUNKS_PER_WARP; ++i) csum[i] = 0.f;
    auto colsum_flush = [&]() {
      if (cs_nt < 0) return;
#pragma unroll
      for (int ci = 0; ci < CHUNKS_PER_WARP; ++ci) {
        const int col = cs_nt * BN + (c_begin + ci) * 32 + lane;
        if (col < p.n) atomicAdd(p.colsum + col, csum[ci]);
        csum[ci] = 0.f;
      }
    };
    for (int w = blockIdx.x; w < p.total_work; w += gridDim.x) {
      const Work wk = decode_work<TN>(p, w);
      if (wk.kb_end <= wk.kb_begin) continue;
      const int j = wk.mt * BM + q * 32 + lane;   // row within the batch (NT) / row of C (TN)
      const bool row_ok = j < p.m;
      // Everything the epilogue needs from global memory is requested BEFORE waiting for the accumulator, so its
      // latency hides behind the MMAs: the bias of this warp's columns (one register per chunk, lane i = column i,
      // broadcast by shuffles later) and the aux row segment of the first chunk.  (Loading them per chunk after the
      // TMEM read cost ~1 500 cycles of exposed latency per chunk and made every K <= 1024 GEMM epilogue-bound.)
      float breg[CHUNKS_PER_WARP];
      uint4 ax[4];
      bool ax_ok = false;
      const __nv_bfloat16* aux_row = nullptr;
      // ReLU-gradient gate as a bit mask written by the forward GEMM: 4 bytes per row and chunk instead of the 64 bytes of
      // the saved activation (32 different rows per warp instruction either way, but 16x fewer bytes and sectors)
      uint32_t gmask[CHUNKS_PER_WARP];
      const long long mrow = static_cast<long long>(wk.b) * p.m + j;
      if constexpr (STORE) {
#pragma unroll
        for (int c = 0; c < CHUNKS_PER_WARP; ++c) {
          const int n0c = wk.nt * BN + (c_begin + c) * 32;
          gmask[c] = (p.gate_mask && row_ok && n0c < p.n) ? __ldg(p.gate_mask + mrow * p.ldmask + (n0c >> 5)) : 0xFFFFFFFFu;
        }
      }
      auto aux_prefetch = [&](int c) {
        const int n0c = wk.nt * BN + c * 32;
        ax_ok = false;
        if (STORE && p.aux_mode && row_ok && n0c + 32 <= p.n) {
          const __nv_bfloat16* ap = aux_row + n0c;
          if ((reinterpret_cast<uintptr_t>(ap) & 15) == 0) {
            ax_ok = true;
#pragma unroll
            for (int i = 0; i < 4; ++i) ax[i] = __ldg(reinterpret_cast<const uint4*>(ap) + i);
          }
        }
      };
      if constexpr (STORE) {
        if constexpr (CSUM) {
          if (wk.nt != cs_nt) {
            colsum_flush();
            cs_nt = wk.nt;
          }
        }
#pragma unroll
        for (int c = 0; c < CHUNKS_PER_WARP; ++c) {
          const int col = wk.nt * BN + (c_begin + c) * 32 + lane;
          breg[c] = (p.bias && col < p.n) ? __ldg(p.bias + col) : 0.f;
        }
        if (p.aux_mode)
          aux_row = p.aux + wk.b * p.aux_batch_stride + static_cast<long long>(j / p.aux_row_div) * p.ldaux;
        aux_prefetch(c_begin);
      }
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;

      if constexpr (STORE) {
        const int fold_rows = p.n_fold > 0 ? p.n / p.n_fold : 1;
#pragma unroll
        for (int ci = 0; ci < CHUNKS_PER_WARP; ++ci) {
          const int c = c_begin + ci;
          const int n0 = wk.nt * BN + c * 32;
          if (n0 >= p.n) break;
          uint32_t v[32];
          tmem_ld32(t_addr + c * 32, v);
          // this chunk's aux values were requested one chunk ago; request the next chunk's before waiting for TMEM
          uint4 cur[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) cur[i] = ax[i];
          const bool cur_ok = ax_ok;
          if (c + 1 < c_end) aux_prefetch(c + 1);
          tmem_ld_wait();
          const bool full_chunk = n0 + 32 <= p.n;
          float f[32];
          const float bv = breg[ci];                   // lane i holds the bias of column n0 + i
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]) + __shfl_sync(0xffffffffu, bv, i);
          if (p.aux_mode && row_ok) {
            float a[32];
            if (cur_ok) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint4 u = cur[i];
                a[i * 8 + 0] = bf16_lo(u.x); a[i * 8 + 1] = bf16_hi(u.x);
                a[i * 8 + 2] = bf16_lo(u.y); a[i * 8 + 3] = bf16_hi(u.y);
                a[i * 8 + 4] = bf16_lo(u.z); a[i * 8 + 5] = bf16_hi(u.z);
                a[i * 8 + 6] = bf16_lo(u.w); a[i * 8 + 7] = bf16_hi(u.w);
              }
            } else {
              const __nv_bfloat16* ap = aux_row + n0;
#pragma unroll
              for (int i = 0; i < 32; ++i) a[i] = (n0 + i < p.n) ? __bfloat162float(ap[i]) : 0.f;
            }
            if (p.aux_mode == 1) {
#pragma unroll
              for (int i = 0; i < 32; ++i) f[i] += a[i];
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) f[i] = a[i] > 0.f ? f[i] : 0.f;
            }
          }
          if (p.gate_mask) {
            const uint32_t mk = gmask[ci];
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = ((mk >> i) & 1u) ? f[i] : 0.f;
          }
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
            if (p.relu_mask && row_ok) {
              uint32_t mk = 0u;
#pragma unroll
              for (int i = 0; i < 32; ++i) mk |= (f[i] > 0.f ? 1u : 0u) << i;
              p.relu_mask[mrow * p.ldmask + (n0 >> 5)] = mk;
            }
          }
          if constexpr (CSUM) {
            // transposing reduction over the warp's 32 rows: at every step a lane keeps the half of its values whose
            // column bit matches its lane bit and adds the partner's copy of them; 31 shuffles leave the sum of
            // column i in lane i
            float r[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = row_ok ? f[i] : 0.f;
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
              const bool upper = (lane & off) != 0;
#pragma unroll
              for (int i = 0; i < off; ++i) {
                const float send = upper ? r[i] : r[i + off];
                const float keep = upper ? r[i + off] : r[i];
                r[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
              }
            }
            csum[ci] += r[0];
          }
          if (p.c_tma == 2) {
            // Wide variant: the warp's chunks leave in PAIRS as one 32-row x 128-byte tile (128-byte swizzle, lane = row)
            // and ONE TMA store per 64 columns - half the store requests of the 64-byte rows below, which bound every
            // GEMM whose K is small against its output (m x 1024 x 256: 345 TFLOP/s = store-engine-bound).  The tile is
            // single-buffered (the staging area is as large as shared memory allows): the previous store of this warp
            // must have READ the tile before the first half of the next pair overwrites it; that wait sits behind this
            // chunk's TMEM read and arithmetic.
            uint8_t* sb = smem + Cfg::STG_OFF + (warp - 4) * 4096;
            const int half = ci & 1;
            if (half == 0) {
              if (lane == 0) tma_store_wait_read<0>();
              __syncwarp();
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4 u;
              u.x = pack_bf16x2(f[i * 8 + 0], f[i * 8 + 1]);
              u.y = pack_bf16x2(f[i * 8 + 2], f[i * 8 + 3]);
              u.z = pack_bf16x2(f[i * 8 + 4], f[i * 8 + 5]);
              u.w = pack_bf16x2(f[i * 8 + 6], f[i * 8 + 7]);
              *reinterpret_cast<uint4*>(sb + lane * 128 + (((half * 4 + i) ^ (lane & 7)) << 4)) = u;
            }
            // the pair is complete after its second half, or after the first if the second lies beyond N (clipped)
            if (half == 1 || n0 + 32 >= p.n) {
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                const int nf = p.n_fold > 0 ? p.n_fold : p.n;
                const int np = n0 - half * 32;         // first column of the pair
                tma_store_4d(&tma_c, sb, np % nf, np / nf, wk.mt * BM + q * 32, wk.b);
                tma_store_commit();
              }
            }
            continue;
          }
          if (p.c_tma) {
            // bf16 tile chunk -> 64-byte-swizzled staging tile (lane = row) -> one TMA store.  Per-lane global stores
            // (32 rows x 16 B per instruction = 32 LSU wavefronts) made the K<=1024 GEMMs epilogue-bound: 1 013 ->
            // 1 370 TFLOP/s without them.
            uint8_t* sb = smem + Cfg::STG_OFF + (warp - 4) * 4096 + stg_buf * 2048;
            if (lane == 0) tma_store_wait_read<1>();   // the store that used this buffer two chunks ago has read it
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4 u;
              u.x = pack_bf16x2(f[i * 8 + 0], f[i * 8 + 1]);
              u.y = pack_bf16x2(f[i * 8 + 2], f[i * 8 + 3]);
              u.z = pack_bf16x2(f[i * 8 + 4], f[i * 8 + 5]);
              u.w = pack_bf16x2(f[i * 8 + 6], f[i * 8 + 7]);
              *reinterpret_cast<uint4*>(sb + lane * 64 + ((i ^ ((lane >> 1) & 3)) << 4)) = u;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              const int nf = p.n_fold > 0 ? p.n_fold : p.n;
              tma_store_4d(&tma_c, sb, n0 % nf, n0 / nf, wk.mt * BM + q * 32, wk.b);
              tma_store_commit();
            }
            stg_buf ^= 1;
            continue;
          }
          if (!row_ok) continue;
          long long out_row = j;
          int out_col = n0;
          if (p.n_fold > 0) {
            out_row = static_cast<long long>(j) * fold_rows + n0 / p.n_fold;
            out_col = n0 % p.n_fold;
          }
          if (p.c_dtype == 0) {
            __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(p.c) + wk.b * p.c_batch_stride + out_row * p.ldc + out_col;
            if (full_chunk && ((reinterpret_cast<uintptr_t>(cp) & 15) == 0)) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                uint4 u;
                u.x = pack_bf16x2(f[i * 8 + 0], f[i * 8 + 1]);
                u.y = pack_bf16x2(f[i * 8 + 2], f[i * 8 + 3]);
                u.z = pack_bf16x2(f[i * 8 + 4], f[i * 8 + 5]);
                u.w = pack_bf16x2(f[i * 8 + 6], f[i * 8 + 7]);
                reinterpret_cast<uint4*>(cp)[i] = u;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (n0 + i < p.n) cp[i] = __float2bfloat16_rn(f[i]);
            }
          } else {
            float* cp = reinterpret_cast<float*>(p.c) + wk.b * p.c_batch_stride + out_row * p.ldc + out_col;
            if (full_chunk && ((reinterpret_cast<uintptr_t>(cp) & 15) == 0)) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                reinterpret_cast<float4*>(cp)[i] = make_float4(f[i * 4], f[i * 4 + 1], f[i * 4 + 2], f[i * 4 + 3]);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (n0 + i < p.n) cp[i] = f[i];
            }
          }
        }
      } else if constexpr (EPI == 2) {
        float* crow = reinterpret_cast<float*>(p.c) + static_cast<long long>(j) * p.ldc;
        for (int c = c_begin; c < c_end; ++c) {
          const int n0 = wk.nt * BN + c * 32;
          if (n0 >= p.n) break;
          uint32_t v[32];
          tmem_ld32(t_addr + c * 32, v);
          tmem_ld_wait();
          if (!row_ok) continue;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (n0 + i < p.n) atomicAdd(crow + n0 + i, __uint_as_float(v[i]));
        }
      } else {
        // log-softmax + NLL family; BN == 256 == N, one thread owns one full row of logits in TMEM
        const long long row = static_cast<long long>(wk.b) * p.m + j;
        const int tgt = row_ok ? static_cast<int>(p.target[row]) : 0;
        float lse, xt = 0.f;
        if (p.nll_mode >= 2 && p.lse != nullptr) {
          // backward with the forward's log-sum-exp at hand: ONE pass over the row instead of three (the logits are
          // recomputed by the same GEMM, so exp(x - lse) is the forward's softmax up to the accumulation order)
          lse = row_ok ? p.lse[row] : 0.f;
        } else {
          float mx = -INFINITY;
          for (int c = 0; c < 8; ++c) {
            uint32_t v[32];
            tmem_ld32(t_addr + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]) + sbias[c * 32 + i]);
          }
          float sum = 0.f;
          for (int c = 0; c < 8; ++c) {
            uint32_t v[32];
            tmem_ld32(t_addr + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float x = __uint_as_float(v[i]) + sbias[c * 32 + i];
              sum += expf(x - mx);
              if (c * 32 + i == tgt) xt = x;
            }
          }
          lse = mx + logf(sum);
        }
        if (p.nll_mode <= 1) {
          if (row_ok) {
            p.lse[row] = lse;
            p.logp_t[row] = xt - lse;
          }
          if (p.nll_mode == 1) {
            for (int c = 0; c < 8; ++c) {
              uint32_t v[32];
              tmem_ld32(t_addr + c * 32, v);
              tmem_ld_wait();
              if (!row_ok) continue;
              float* op = p.logp + row * p.ldlogp + c * 32;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float4 o;
                o.x = __uint_as_float(v[i * 4 + 0]) + sbias[c * 32 + i * 4 + 0] - lse;
                o.y = __uint_as_float(v[i * 4 + 1]) + sbias[c * 32 + i * 4 + 1] - lse;
                o.z = __uint_as_float(v[i * 4 + 2]) + sbias[c * 32 + i * 4 + 2] - lse;
                o.w = __uint_as_float(v[i * 4 + 3]) + sbias[c * 32 + i * 4 + 3] - lse;
                reinterpret_cast<float4*>(op)[i] = o;
              }
            }
          }
        } else {
          float rg = 0.f, gsum = 0.f;
          if (row_ok) {
            if (p.nll_mode == 2) {
              rg = p.row_grad[row];
            } else {
              const float4* gp = reinterpret_cast<const float4*>(p.g + row * p.ldg);
              for (int i = 0; i < 64; ++i) {
                const float4 t = __ldg(gp + i);
                gsum += (t.x + t.y) + (t.z + t.w);
              }
            }
          }
          for (int c = 0; c < 8; ++c) {
            uint32_t v[32];
            tmem_ld32(t_addr + c * 32, v);
            tmem_ld_wait();
            if (!row_ok) continue;
            float d[32];
            if (p.nll_mode == 2) {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float pr = expf(__uint_as_float(v[i]) + sbias[c * 32 + i] - lse);
                d[i] = rg * ((c * 32 + i == tgt ? 1.f : 0.f) - pr);
              }
            } else {
              const float4* gp = reinterpret_cast<const float4*>(p.g + row * p.ldg + c * 32);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 t = __ldg(gp + i);
                const float gv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float pr = expf(__uint_as_float(v[i * 4 + e]) + sbias[c * 32 + i * 4 + e] - lse);
                  d[i * 4 + e] = gv[e] - pr * gsum;
                }
              }
            }
            __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(p.c) + row * p.ldc + c * 32;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4 u;
              u.x = pack_bf16x2(d[i * 8 + 0], d[i * 8 + 1]);
              u.y = pack_bf16x2(d[i * 8 + 2], d[i * 8 + 3]);
              u.z = pack_bf16x2(d[i * 8 + 4], d[i * 8 + 5]);
              u.w = pack_bf16x2(d[i * 8 + 6], d[i * 8 + 7]);
              reinterpret_cast<uint4*>(dp)[i] = u;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if constexpr (CSUM) colsum_flush();
    if (STORE && p.c_tma && lane == 0) tma_store_wait_read<0>();   // staging tiles stay valid until read
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// Small-M NT GEMM (M <= 512: the per-sample GEMMs of generation, model.py:289-351, and tiny shapes).
// With so few rows the 128 x 256 persistent kernel runs on a handful of SMs, each streaming its
// whole K extent through one SM's L2 port with only 4 x 48 KB in flight (256 x 1024 x 1024: 8 CTAs,
// 14 us).  Here the tile is 128 x 64 and the K extent is SPLIT over a thread-block cluster of CS CTAs
// (2 x 16 tiles x 4 = 128 CTAs for the same shape, every CTA with its whole K share in flight at once).
// Reduction, push model (as in the recurrent kernel): every rank sends, for each destination rank, the
// 64/CS columns that rank finalises straight into the destination's shared memory with st.async; the
// bytes complete on the destination's mbarrier.  The destination then sums the CS slices in a fixed order
// from its own shared memory, applies bias / aux / ReLU and stores.  Deterministic, no atomics, no
// workspace.  (A pull model - cluster barrier, then DSMEM loads of the peers' partial tiles - was measured
// 3 us slower per launch: the 16-byte remote loads are latency- and request-rate-bound.)
// ---------------------------------------------------------------------------------------------
constexpr int SBN = 64;
constexpr int S_STAGES = 4;
constexpr int S_STAGE = BM * BK * 2 + SBN * BK * 2;      // 24 KB
constexpr int S_RECV_OFF = S_STAGES * S_STAGE;           // recv[src rank][128 rows][64/CS columns] fp32 = 32 KB
constexpr int S_BAR_OFF = S_RECV_OFF + BM * SBN * 4;
constexpr int S_TOTAL = S_BAR_OFF + 256 + 1024;

#ifdef SRNN_SMALL_TS
__device__ unsigned long long g_small_ts[512 * 8];
#define SMALL_TS(slot_)                                                                          \
  do {                                                                                           \
    if (threadIdx.x == (slot_ == 2 ? 32 : (slot_ >= 3 ? 128 : 0)) && blockIdx.x < 512) {         \
      unsigned long long t_;                                                                     \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                     \
      g_small_ts[blockIdx.x * 8 + (slot_)] = t_;                                                 \
    }                                                                                            \
  } while (0)
#else
#define SMALL_TS(slot_)
#endif

template <int CS>
__global__ void __launch_bounds__(384, 1)
gemm_small_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                  const GemmParams p) {
  constexpr uint32_t IDESC = idesc_bf16(BM, SBN, false, false);
  constexpr int CPC = SBN / CS;                        // columns finalised by one rank
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* recv = reinterpret_cast<float*>(smem + S_RECV_OFF);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S_BAR_OFF);
  uint64_t* empty = full + S_STAGES;
  uint64_t* tfull = empty + S_STAGES;
  uint64_t* recv_bar = tfull + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(recv_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  SMALL_TS(0);
  pdl_launch_dependents();                             // the next kernel's launch may overlap this one (it waits itself)
  const uint32_t crank = CS > 1 ? cluster_ctarank() : 0u;
  const int tile = blockIdx.x / CS;
  const int nt = tile % p.tiles_n, mt = tile / p.tiles_n;
  const int kb_begin = static_cast<int>(crank) * p.kb_per_split;
  const int kb_end = min(kb_begin + p.kb_per_split, p.total_kb);
  const int nkb = max(kb_end - kb_begin, 0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tfull, 1);
    mbar_init(recv_bar, 1);
    fence_barrier_init();
    if (CS > 1) mbar_expect_tx(recv_bar, static_cast<uint32_t>((CS - 1) * BM * CPC * 4));   // the peers' slices
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, SBN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (CS > 1) cluster_arrive();                        // (waited for just before the first remote push)
  const uint32_t tmem_base = *tmem_slot;
  SMALL_TS(1);
  pdl_wait();                                          // set-up above overlapped the previous kernel; its data from here on

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sa = smem + stage * S_STAGE;
        mbar_expect_tx(&full[stage], S_STAGE);
        tma_load_3d(sa, &tma_a, &full[stage], kb * BK, mt * BM, 0);
        tma_load_2d(sa + BM * BK * 2, &tma_b, &full[stage], kb * BK, nt * SBN);
        if (++stage == S_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && nkb > 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&full[stage], phase);
        if (kb == kb_begin) SMALL_TS(2);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + stage * S_STAGE);
        const uint32_t b_addr = a_addr + BM * BK * 2;
#pragma unroll
        for (int k16 = 0; k16 < BK / 16; ++k16)
          umma_bf16(tmem_base, smem_desc_sw128(a_addr + k16 * 32, 16, 1024), smem_desc_sw128(b_addr + k16 * 32, 16, 1024),
                    IDESC, (kb > kb_begin || k16 > 0) ? 1u : 0u);
        umma_commit(&empty[stage]);
        if (++stage == S_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(tfull);
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ---- partial tile out of TMEM; each group of 4 columns goes to the rank that finalises it -------------
    const int q = warp & 3, half = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    uint32_t v[32];
    if (nkb > 0) {
      mbar_wait(tfull, 0);
      SMALL_TS(3);
      tc_fence_after();
      tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + half * 32, v);
      tmem_ld_wait();
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = 0u;          // this rank got no K blocks
    }
    if (CS > 1) cluster_wait();                        // every rank's barriers are initialised
    const uint32_t recv_addr = smem_u32(recv);
    const uint32_t bar_addr = smem_u32(recv_bar);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int col = half * 32 + i * 4;               // tile column of this float4
      const int dst = col / CPC;
      const uint32_t off = static_cast<uint32_t>(((static_cast<int>(crank) * BM + row) * CPC + col % CPC) * 4);
      if (CS == 1 || dst == static_cast<int>(crank)) {
        *reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(recv) + off) =
            make_float4(__uint_as_float(v[i * 4]), __uint_as_float(v[i * 4 + 1]), __uint_as_float(v[i * 4 + 2]),
                        __uint_as_float(v[i * 4 + 3]));
      } else {
        st_async_v4(mapa(recv_addr + off, static_cast<uint32_t>(dst)), __uint_as_float(v[i * 4]),
                    __uint_as_float(v[i * 4 + 1]), __uint_as_float(v[i * 4 + 2]), __uint_as_float(v[i * 4 + 3]),
                    mapa(bar_addr, static_cast<uint32_t>(dst)));
      }
    }
    tc_fence_before();
    asm volatile("bar.sync 1, 256;" ::: "memory");     // this rank's own slice is complete in shared memory
    if (CS > 1) mbar_wait(recv_bar, 0);                // ... and so are the peers' slices
    SMALL_TS(4);

    // ---- sum the CS slices of this rank's columns (fixed order), epilogue, store ---------------------------
    constexpr int CPT = CPC / 2;                       // columns per thread: 32 / 16 / 8
    const int tid = threadIdx.x - 128;
    const int r2 = tid >> 1;
    const int c0 = (tid & 1) * CPT;
    float f[CPT];
#pragma unroll
    for (int i = 0; i < CPT; ++i) f[i] = 0.f;
#pragma unroll
    for (int src = 0; src < CS; ++src) {
      const float4* sp = reinterpret_cast<const float4*>(recv + (src * BM + r2) * CPC + c0);
#pragma unroll
      for (int i = 0; i < CPT / 4; ++i) {
        const float4 t = sp[i];
        f[i * 4 + 0] += t.x; f[i * 4 + 1] += t.y; f[i * 4 + 2] += t.z; f[i * 4 + 3] += t.w;
      }
    }
    const int gr = mt * BM + r2;
    const int gc = nt * SBN + static_cast<int>(crank) * CPC + c0;
    if (gr < p.m && gc < p.n) {
      const bool full_cols = gc + CPT <= p.n;
      if (p.bias) {
#pragma unroll
        for (int i = 0; i < CPT; ++i)
          if (full_cols || gc + i < p.n) f[i] += __ldg(p.bias + gc + i);
      }
      if (p.aux_mode) {
        const __nv_bfloat16* ap = p.aux + static_cast<long long>(gr / p.aux_row_div) * p.ldaux + gc;
#pragma unroll
        for (int i = 0; i < CPT; ++i)
          if (full_cols || gc + i < p.n) f[i] += __bfloat162float(ap[i]);
      }
      if (p.relu) {
#pragma unroll
        for (int i = 0; i < CPT; ++i) f[i] = fmaxf(f[i], 0.f);
      }
      if (p.c_dtype == 0) {
        __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(p.c) + static_cast<long long>(gr) * p.ldc + gc;
        if (full_cols && ((reinterpret_cast<uintptr_t>(cp) & 15) == 0)) {
#pragma unroll
          for (int i = 0; i < CPT / 8; ++i) {
            uint4 u;
            u.x = pack_bf16x2(f[i * 8 + 0], f[i * 8 + 1]);
            u.y = pack_bf16x2(f[i * 8 + 2], f[i * 8 + 3]);
            u.z = pack_bf16x2(f[i * 8 + 4], f[i * 8 + 5]);
            u.w = pack_bf16x2(f[i * 8 + 6], f[i * 8 + 7]);
            reinterpret_cast<uint4*>(cp)[i] = u;
          }
        } else {
#pragma unroll
          for (int i = 0; i < CPT; ++i)
            if (gc + i < p.n) cp[i] = __float2bfloat16_rn(f[i]);
        }
      } else {
        float* cp = reinterpret_cast<float*>(p.c) + static_cast<long long>(gr) * p.ldc + gc;
        if (full_cols && ((reinterpret_cast<uintptr_t>(cp) & 15) == 0)) {
#pragma unroll
          for (int i = 0; i < CPT / 4; ++i)
            reinterpret_cast<float4*>(cp)[i] = make_float4(f[i * 4], f[i * 4 + 1], f[i * 4 + 2], f[i * 4 + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < CPT; ++i)
            if (gc + i < p.n) cp[i] = f[i];
        }
      }
    }
    SMALL_TS(5);
  }
  // A rank leaves only after every slice addressed to it has landed (recv_bar above); its own pushes carry their
  // data with them.  Warps 0-3 still owe the wait half of the cluster barrier.
  if (CS > 1 && warp < 4) cluster_wait();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, SBN);
  SMALL_TS(6);
}

#ifdef SRNN_SMALL_TS
extern "C" int srnn_debug_small_ts(unsigned long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, g_small_ts, sizeof(g_small_ts)) == cudaSuccess ? 0 : 1;
}
#endif

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <int BN, bool TN, int EPI>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& ta2,
                  const GemmParams& p, cudaStream_t stream) {
  auto kern = gemm_kernel<BN, TN, EPI>;
  static DeviceOnce configured;     // per instantiation and device
  if (!configured.test()) {
    SRNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemCfg<BN>::TOTAL));
    configured.set();
  }
  int cap = sm_count();
  if (p.max_ctas > 0 && p.max_ctas < cap) cap = p.max_ctas;    // leave SMs to a concurrently running kernel
  int grid = p.total_work < cap ? p.total_work : cap;
  if (grid < 1) return SRNN_OK;
  kern<<<grid, gemm_threads<EPI>(), SmemCfg<BN>::TOTAL, stream>>>(ta, tb, tc, ta2, p);
  SRNN_CUDA(cudaGetLastError());
  return SRNN_OK;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int run_nt(const srnn_gemm_args* a, GemmParams& p, bool nll, cudaStream_t stream) {
  SRNN_CHECK_ARG(a->lda % 8 == 0 && a->ldb % 8 == 0 && a->a_batch_stride % 8 == 0,
                 "gemm NT: lda/ldb/a_batch_stride must be multiples of 8 elements (lda=%lld ldb=%lld abs=%lld)",
                 (long long)a->lda, (long long)a->ldb, (long long)a->a_batch_stride);
  SRNN_CHECK_ARG(aligned16(a->a) && aligned16(a->b), "gemm NT: operand pointers must be 16-byte aligned");
  const int bn = (nll || a->n % 256 == 0 || a->n > 512) ? 256 : 128;
  p.tiles_m = (a->m + BM - 1) / BM;
  p.tiles_n = (a->n + bn - 1) / bn;
  p.kb_per_batch = (a->k + BK - 1) / BK;
  p.splits = 1;
  p.kb_per_split = p.kb_per_batch;
  p.total_kb = p.kb_per_batch;
  p.total_work = a->batch * p.tiles_m * p.tiles_n;
  CUtensorMap ta, tb, ta2;
  const int k_a1 = a->a2 ? a->k1 : a->k;               // columns of the first A operand
  p.kb_a1 = a->a2 ? a->k1 / BK : p.kb_per_batch;
  {
    const uint64_t dims[3] = {(uint64_t)k_a1, (uint64_t)a->m, (uint64_t)a->batch};
    const uint64_t bs = a->batch > 1 ? (uint64_t)a->a_batch_stride : (uint64_t)a->lda * (uint64_t)a->m;
    const uint64_t strides[2] = {(uint64_t)a->lda * 2, bs * 2};
    const uint32_t box[3] = {BK, BM, 1};
    int rc = make_tmap_bf16(&ta, a->a, 3, dims, strides, box, true);
    if (rc) return rc;
  }
  ta2 = ta;
  if (a->a2) {
    SRNN_CHECK_ARG(a->k1 > 0 && a->k1 < a->k && a->k1 % BK == 0, "gemm NT: k1 must be a multiple of 64 inside (0, k)");
    SRNN_CHECK_ARG(a->lda2 % 8 == 0 && a->a2_batch_stride % 8 == 0 && aligned16(a->a2),
                   "gemm NT: second A operand must be 16-byte aligned with strides in multiples of 8 elements");
    const uint64_t dims[3] = {(uint64_t)(a->k - a->k1), (uint64_t)a->m, (uint64_t)a->batch};
    const uint64_t bs = a->batch > 1 ? (uint64_t)a->a2_batch_stride : (uint64_t)a->lda2 * (uint64_t)a->m;
    const uint64_t strides[2] = {(uint64_t)a->lda2 * 2, bs * 2};
    const uint32_t box[3] = {BK, BM, 1};
    int rc = make_tmap_bf16(&ta2, a->a2, 3, dims, strides, box, true);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)a->k, (uint64_t)a->n};
    const uint64_t strides[1] = {(uint64_t)a->ldb * 2};
    const uint32_t box[2] = {BK, (uint32_t)bn};
    int rc = make_tmap_bf16(&tb, a->b, 2, dims, strides, box, true);
    if (rc) return rc;
  }
  if (nll) return launch<256, false, 1>(ta, tb, ta, ta, p, stream);
  // bf16 outputs with 16-byte-aligned rows leave through shared memory and TMA stores; C is described as
  // [batch][m][n / n_fold][n_fold] so that the folded (upsampling) layout is the same code path
  CUtensorMap tc = ta;
  p.c_tma = 0;
  if (a->c_dtype == 0 && a->ldc % 8 == 0 && a->c_batch_stride % 8 == 0 && aligned16(a->c)) {
    const uint64_t nf = a->n_fold > 0 ? (uint64_t)a->n_fold : (uint64_t)a->n;
    const uint64_t fold_rows = (uint64_t)a->n / nf;
    const uint64_t row_bytes = (uint64_t)a->ldc * 2;
    const uint64_t bs = a->batch > 1 ? (uint64_t)a->c_batch_stride * 2 : row_bytes * fold_rows * (uint64_t)a->m;
    const uint64_t dims[4] = {nf, fold_rows, (uint64_t)a->m, (uint64_t)a->batch};
    const uint64_t strides[3] = {row_bytes, row_bytes * fold_rows, bs};
    // 64-column (128-byte) store tiles unless a fold boundary could fall inside one
    static const bool narrow_only = getenv("SRNN_GEMM_STORE64") != nullptr;      // A/B switch for measurements
    const bool wide = !narrow_only && (a->n_fold == 0 || a->n_fold % 64 == 0);
    int rc;
    if (wide) {
      const uint32_t box[4] = {64, 1, 32, 1};
      rc = make_tmap_bf16(&tc, a->c, 4, dims, strides, box, true);
    } else {
      const uint32_t box[4] = {32, 1, 32, 1};
      rc = make_tmap_bf16_sw64(&tc, a->c, 4, dims, strides, box);
    }
    if (rc) return rc;
    p.c_tma = wide ? 2 : 1;
  }
  if (p.colsum) {
    if (bn == 256) return launch<256, false, 3>(ta, tb, tc, ta2, p, stream);
    return launch<128, false, 3>(ta, tb, tc, ta2, p, stream);
  }
  if (bn == 256) return launch<256, false, 0>(ta, tb, tc, ta2, p, stream);
  return launch<128, false, 0>(ta, tb, tc, ta2, p, stream);
}

template <int CS>
static int launch_small(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t stream) {
  auto kern = gemm_small_kernel<CS>;
  static DeviceOnce configured;     // per instantiation and device
  if (!configured.test()) {
    SRNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S_TOTAL));
    configured.set();
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(p.tiles_m * p.tiles_n * CS);
  cfg.blockDim = dim3(384);
  cfg.dynamicSmemBytes = S_TOTAL;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CS > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = CS;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  SRNN_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, p));
  return SRNN_OK;
}

// M <= 512, one batch, plain epilogue: 128 x 64 tiles, K split over a cluster while the CTAs still fit one wave
static int run_nt_small(const srnn_gemm_args* a, GemmParams& p, cudaStream_t stream) {
  SRNN_CHECK_ARG(a->lda % 8 == 0 && a->ldb % 8 == 0, "gemm NT: lda/ldb must be multiples of 8 elements (lda=%lld ldb=%lld)",
                 (long long)a->lda, (long long)a->ldb);
  SRNN_CHECK_ARG(aligned16(a->a) && aligned16(a->b), "gemm NT: operand pointers must be 16-byte aligned");
  p.tiles_m = (a->m + BM - 1) / BM;
  p.tiles_n = (a->n + SBN - 1) / SBN;
  p.total_kb = p.kb_per_batch = (a->k + BK - 1) / BK;
  const int tiles = p.tiles_m * p.tiles_n;
  int cs = 1;
  while (cs < 4 && tiles * cs * 2 <= sm_count() && p.total_kb >= cs * 4) cs *= 2;   // one wave of CTAs, >= 2 K blocks each
#ifdef SRNN_SMALL_TS
  if (const char* e = getenv("SRNN_SMALL_CS")) cs = atoi(e);
#endif
  p.splits = cs;
  p.kb_per_split = (p.total_kb + cs - 1) / cs;
  p.total_work = tiles;
  CUtensorMap ta, tb;
  {
    const uint64_t dims[3] = {(uint64_t)a->k, (uint64_t)a->m, 1};
    const uint64_t strides[2] = {(uint64_t)a->lda * 2, (uint64_t)a->lda * 2 * (uint64_t)a->m};
    const uint32_t box[3] = {BK, BM, 1};
    int rc = make_tmap_bf16(&ta, a->a, 3, dims, strides, box, true);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)a->k, (uint64_t)a->n};
    const uint64_t strides[1] = {(uint64_t)a->ldb * 2};
    const uint32_t box[2] = {BK, SBN};
    int rc = make_tmap_bf16(&tb, a->b, 2, dims, strides, box, true);
    if (rc) return rc;
  }
  switch (cs) {
    case 4: return launch_small<4>(ta, tb, p, stream);
    case 2: return launch_small<2>(ta, tb, p, stream);
    default: return launch_small<1>(ta, tb, p, stream);
  }
}

static int run_tn(const srnn_gemm_args* a, GemmParams& p, cudaStream_t stream) {
  SRNN_CHECK_ARG(a->lda % 8 == 0 && a->ldb % 8 == 0 && a->a_batch_stride % 8 == 0 && a->b_batch_stride % 8 == 0,
                 "gemm TN: lda/ldb/batch strides must be multiples of 8 elements");
  SRNN_CHECK_ARG(aligned16(a->a) && aligned16(a->b), "gemm TN: operand pointers must be 16-byte aligned");
  SRNN_CHECK_ARG(a->c_dtype == 1 && a->n_fold == 0 && !a->bias && !a->aux && !a->relu,
                 "gemm TN: output is fp32 atomic-accumulate only, no epilogue options");
  const int bn = (a->n % 256 == 0 || a->n > 512) ? 256 : 128;
  p.tiles_m = (a->m + BM - 1) / BM;
  p.tiles_n = (a->n + bn - 1) / bn;
  p.kb_per_batch = (a->k + BK - 1) / BK;
  p.total_kb = p.kb_per_batch * a->batch;
  const int tiles = p.tiles_m * p.tiles_n;
  // split-K factor: the persistent grid runs ceil(items / SMs) rounds, each costing the K-blocks of one
  // split (4 MMAs x 128 cycles at BN=256) plus a full-tile fp32 atomic epilogue (contended when many
  // splits hit the same tile).  Minimise rounds x (K-blocks per split x 512 + 32768) cycles: e.g. 32 tiles
  // -> 9 splits = 288 items = 2 full rounds instead of 10 splits = 320 items = 3 rounds.
  int sms = sm_count();
  if (p.max_ctas > 0 && p.max_ctas < sms) sms = p.max_ctas;
  int best = 1;
  double best_cost = 1e30;
  for (int sp = 1; sp <= 64 && sp <= p.total_kb; ++sp) {
    const long long items = static_cast<long long>(tiles) * sp;
    const long long rounds = (items + sms - 1) / sms;
    const long long kb = (p.total_kb + sp - 1) / sp;
    const double cost = static_cast<double>(rounds) * (static_cast<double>(kb) * 512.0 + 32768.0);
    if (cost < best_cost) {
      best_cost = cost;
      best = sp;
    }
  }
  p.kb_per_split = (p.total_kb + best - 1) / best;
  p.splits = (p.total_kb + p.kb_per_split - 1) / p.kb_per_split;
  p.total_work = tiles * p.splits;
  CUtensorMap ta, tb;
  // MN-major operands: smem wants [MN block of 64][k row][64 elements = 128 B].  When M and N are whole numbers of
  // 64-element blocks the matrix is viewed as [batch][MN/64][k][64] (non-monotonic strides) and ONE box brings all
  // blocks of a tile; otherwise one box per 64-element block.
  p.tn_4d = (a->m % 64 == 0 && a->n % 64 == 0) ? 1 : 0;
  {
    const uint64_t rows = (uint64_t)(a->a_row_offset + a->k);
    const uint64_t bs = a->batch > 1 ? (uint64_t)a->a_batch_stride : (uint64_t)a->lda * rows;
    int rc;
    if (p.tn_4d) {
      const uint64_t dims[4] = {64, rows, (uint64_t)a->m / 64, (uint64_t)a->batch};
      const uint64_t strides[3] = {(uint64_t)a->lda * 2, 128, bs * 2};
      const uint32_t box[4] = {64, 64, BM / 64, 1};
      rc = make_tmap_bf16(&ta, a->a, 4, dims, strides, box, true);
    } else {
      const uint64_t dims[3] = {(uint64_t)a->m, rows, (uint64_t)a->batch};
      const uint64_t strides[2] = {(uint64_t)a->lda * 2, bs * 2};
      const uint32_t box[3] = {64, 64, 1};
      rc = make_tmap_bf16(&ta, a->a, 3, dims, strides, box, true);
    }
    if (rc) return rc;
  }
  {
    const uint64_t rows = (uint64_t)(a->b_row_offset + a->k);
    const uint64_t bs = a->batch > 1 ? (uint64_t)a->b_batch_stride : (uint64_t)a->ldb * rows;
    int rc;
    if (p.tn_4d) {
      const uint64_t dims[4] = {64, rows, (uint64_t)a->n / 64, (uint64_t)a->batch};
      const uint64_t strides[3] = {(uint64_t)a->ldb * 2, 128, bs * 2};
      const uint32_t box[4] = {64, 64, (uint32_t)bn / 64, 1};
      rc = make_tmap_bf16(&tb, a->b, 4, dims, strides, box, true);
    } else {
      const uint64_t dims[3] = {(uint64_t)a->n, rows, (uint64_t)a->batch};
      const uint64_t strides[2] = {(uint64_t)a->ldb * 2, bs * 2};
      const uint32_t box[3] = {64, 64, 1};
      rc = make_tmap_bf16(&tb, a->b, 3, dims, strides, box, true);
    }
    if (rc) return rc;
  }
  if (bn == 256) return launch<256, true, 2>(ta, tb, ta, ta, p, stream);
  return launch<128, true, 2>(ta, tb, ta, ta, p, stream);
}

}  // namespace srnn

using namespace srnn;

extern "C" int srnn_gemm_bf16(const srnn_gemm_args* a, srnn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SRNN_CHECK_ARG(a != nullptr, "gemm: null args");
  SRNN_CHECK_ARG(a->m > 0 && a->n > 0 && a->k > 0 && a->batch > 0, "gemm: m,n,k,batch must be positive (%d %d %d %d)",
                 a->m, a->n, a->k, a->batch);
  SRNN_CHECK_ARG(a->a && a->b && a->c, "gemm: null operand");
  GemmParams p{};
  p.m = a->m; p.n = a->n; p.k = a->k; p.batch = a->batch;
  p.a_row_offset = a->a_row_offset; p.b_row_offset = a->b_row_offset;
  p.c = a->c; p.ldc = a->ldc; p.c_batch_stride = a->c_batch_stride;
  p.c_dtype = a->c_dtype; p.n_fold = a->n_fold;
  p.bias = a->bias;
  p.aux = static_cast<const __nv_bfloat16*>(a->aux); p.ldaux = a->ldaux; p.aux_batch_stride = a->aux_batch_stride;
  p.aux_mode = a->aux ? a->aux_mode : 0; p.relu = a->relu;
  p.aux_row_div = a->aux_row_div > 0 ? a->aux_row_div : 1;
  p.max_ctas = a->max_ctas;
  p.colsum = a->op == 0 ? a->colsum : nullptr;
  p.relu_mask = a->op == 0 ? a->relu_mask : nullptr;
  p.gate_mask = a->op == 0 ? a->gate_mask : nullptr;
  p.ldmask = a->ldmask;
  SRNN_CHECK_ARG((!a->relu_mask && !a->gate_mask) || (a->op == 0 && a->ldmask * 32 >= a->n && a->n_fold == 0),
                 "gemm: relu_mask / gate_mask need op NT, no n_fold and ldmask >= ceil(n / 32)");
  SRNN_CHECK_ARG(!a->relu_mask || a->relu, "gemm: relu_mask is an output of the ReLU epilogue");
  if (a->op == 0) {
    SRNN_CHECK_ARG(a->n_fold == 0 || (a->n % a->n_fold == 0 && a->n_fold % 32 == 0 && !a->aux),
                   "gemm NT: n_fold must divide n, be a multiple of 32, and exclude aux");
    if (a->m <= 512 && a->batch == 1 && a->n_fold == 0 && p.aux_mode != 2 && a->max_ctas == 0 && !a->colsum && !a->a2 &&
        !a->relu_mask && !a->gate_mask)
      return run_nt_small(a, p, stream);
    return run_nt(a, p, false, stream);
  }
  SRNN_CHECK_ARG(a->op == 1, "gemm: op must be 0 (NT) or 1 (TN)");
  SRNN_CHECK_ARG(!a->colsum, "gemm TN: colsum is an NT epilogue option");
  return run_tn(a, p, stream);
}

extern "C" int srnn_gemm_nll(const srnn_nll_args* a, srnn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SRNN_CHECK_ARG(a != nullptr && a->m > 0 && a->k > 0, "gemm_nll: bad shape");
  SRNN_CHECK_ARG(a->a && a->w && a->target, "gemm_nll: null operand");
  SRNN_CHECK_ARG(a->mode >= 0 && a->mode <= 3, "gemm_nll: mode must be 0..3");
  if (a->mode <= 1) SRNN_CHECK_ARG(a->lse && a->logp_target, "gemm_nll: lse/logp_target required");
  if (a->mode == 1) SRNN_CHECK_ARG(a->logp && a->ldlogp % 4 == 0, "gemm_nll: logp (ld %% 4 == 0) required for mode 1");
  if (a->mode == 2) SRNN_CHECK_ARG(a->row_grad && a->dlogits, "gemm_nll: row_grad/dlogits required for mode 2");
  if (a->mode == 3) SRNN_CHECK_ARG(a->g && a->dlogits && a->ldg % 4 == 0, "gemm_nll: g/dlogits required for mode 3");
  if (a->mode >= 2) SRNN_CHECK_ARG(a->lddlogits % 8 == 0, "gemm_nll: lddlogits must be a multiple of 8");
  srnn_gemm_args g{};
  g.op = 0; g.m = a->m; g.n = 256; g.k = a->k; g.batch = 1;
  g.a = a->a; g.lda = a->lda; g.b = a->w; g.ldb = a->ldw;
  GemmParams p{};
  p.m = a->m; p.n = 256; p.k = a->k; p.batch = 1;
  p.bias = a->bias;
  p.nll_mode = a->mode; p.target = a->target; p.lse = a->lse; p.logp_t = a->logp_target;
  p.logp = a->logp; p.ldlogp = a->ldlogp; p.row_grad = a->row_grad; p.g = a->g; p.ldg = a->ldg;
  p.c = a->dlogits; p.ldc = a->lddlogits;
  return run_nt(&g, p, true, stream);
}
