// Persistent recurrent kernels for sm_100a: GRU (replaces torch.nn.GRU, model.py:110,152, and its
// autograd backward) and an LSTM variant (BASELINE config 3 extension).
//
// One cooperative launch runs all T timesteps of a tier; the recurrent weights never leave shared
// memory.  Work decomposition (H hidden units, batch rows M <= 64, K = H forward / 3H backward):
//
//   * the grid is H/U CTAs (U = 8 units per CTA, optionally 16) in thread-block clusters of C CTAs
//     (forward at H = 1024: C = 1, every CTA holds its 24 gate rows of W_hh for the whole K; backward: C = 4);
//   * a cluster owns U*C hidden units: forward its 3*U*C gate rows of W_hh, backward its U*C rows of
//     W_hh^T.  Inside the cluster the reduction dimension is SPLIT: CTA rank c keeps only the K/C
//     slice [c*K/C, (c+1)*K/C) of those rows resident (48 KB at H=1024, 128-byte-swizzled K-major);
//   * per timestep each CTA
//       1. polls a grid-wide arrival counter in L2 (relaxed loads) until every CTA has published its
//          part of h_{t-1} (backward: of the gate gradients of step t+1), then a generic->async proxy fence,
//       2. TMA-loads only its K slice of that [batch, K] matrix (ONE 4-D box, one mbarrier),
//       3. multiplies it with tcgen05.mma (M=64, N=U*C*{3,1}), issued by four threads into four partial
//          accumulators in TMEM,
//       4. (C > 1) reads the partial tile from TMEM and PUSHES, for every peer rank, the columns of that rank's
//          units into the peer's shared memory with st.async; the bytes complete on the peer's mbarrier,
//       5. sums the partials, VALIDATES the operand it read (a bf16 NaN sentinel in any element that had not
//          arrived yet makes the row's accumulators NaN: the attempt is then repeated), finishes the gate math in
//          fp32 registers (MUFU tanh), stores its U columns of h_t (bf16, time-major exchange buffer) and arrives
//          on the counter with a RELAXED increment; the batch-major copy of h_t and the saved gates are written
//          after the arrival, off the critical path.
//
// How it got here (B=64, H=1024, us per step fwd/bwd; profiles/r01_gru_*, r02_gru_*): one CTA per 8 units doing the
// whole K loop 9.5/18.0 (tensor pipe 5 % busy, L2 5 % busy: single-thread issue + handshake latency) ->
// cluster split-K with smem staging + DSMEM loads 6.8/7.5 -> st.async push exchange 5.4/6.0 -> MMA issuers with
// separate partial accumulators + running descriptors 5.0/5.35 -> forward without clusters 3.97/4.6 -> release /
// acquire fences replaced by the validated read (same box: 4.77/5.57 strict -> 3.54/4.26).
//
// The fp32 recurrent state of a unit never leaves the registers of its owner thread.
#include <mutex>

#include "common.cuh"

namespace srnn {

// warps 0, 1 and 6.. MMA issuers (MW of them: a tcgen05.mma of this size costs ~90 cycles of issue from one
// thread, ~24 of tensor pipe; thread 0 is also the TMA producer), warps 2-5 epilogue
constexpr int gru_threads(int mw) { return 32 * (4 + mw); }
// hidden units finalised per CTA: 8 (H/8 CTAs, the fastest step) or 16 (H/16 CTAs, which leaves more than
// half of the SMs free for GEMMs running concurrently on another stream)
constexpr int GRU_M = 64;          // batch rows per launch (MMA M)
constexpr int GRU_SLOT = GRU_M * 128;   // one [64 rows][64 bf16] K block
constexpr int GRU_MAX_KBC = 32;         // K blocks of the exchanged operand per CTA (one landing barrier each)

struct GruParams {
  int batch, steps, hidden, ext_batch;
  int kbc;                         // K blocks (of 64) per CTA
  int hslot;                       // bytes between K blocks of the staged [batch, K slice] operand
  int one_box;                     // the whole slice arrives as ONE 4-D TMA box (else one box per K block)
  int groups;                      // > 1: single-timestep forward over `groups` blocks of 64 rows in ONE launch (generation)
  int box_rows;                    // rows of one TMA box of the exchanged operand (min(batch, group_rows))
  int group_rows;                  // batch rows per group: 64 (one full MMA tile), or 32 (tuning flag 1 << 27, batch <= 64)
  int group_sync;                  // groups > 1: every group has its own arrival counter, published and awaited per group
  const __nv_bfloat16* gi;
  const float* b_hh;
  __nv_bfloat16* h_ext;
  __nv_bfloat16* hall;
  float* h_state;
  __nv_bfloat16* gates;
  const __nv_bfloat16* dh_out;
  __nv_bfloat16* dgi;
  __nv_bfloat16* dgh;
  float* dh0;
  float* c_state;                  // LSTM: fp32 [batch, H] cell state, in = c_init, out = c_T
  const float* c_init;             // LSTM backward: the cell state the forward started from
  float* dc0;                      // LSTM backward: dL/dc_init
  float* db_ih;                    // backward (nullable): += column sums of dgi
  float* db_hh;                    // backward (nullable): += column sums of dgh
  uint32_t* sync;
  int flags;
  unsigned long long* ts;          // debug timestamps [256][8] of CTA 0 (nullable)
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// debug stamps: clock64 of CTA 0 per timestep [step][8], or (flag 128) the global timer of EVERY CTA at
// timestep 24 [cta][8] to see the skew between CTAs
#define GRU_TS(slot_, step_)                                                                 \
  do {                                                                                       \
    if (p.ts) {                                                                              \
      if (p.flags & 128) {                                                                   \
        if ((step_) == 24 && blockIdx.x < 256) p.ts[blockIdx.x * 8 + (slot_)] = globaltimer_ns(); \
      } else if (blockIdx.x == 0 && (step_) < 256) {                                         \
        p.ts[(step_) * 8 + (slot_)] = clock64();                                             \
      }                                                                                      \
    }                                                                                        \
  } while (0)

// MUFU approximations (max relative error 2^-11, below the bf16 rounding of the matmul operand)
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }

template <int U>
__device__ __forceinline__ void load_units(const __nv_bfloat16* p, float (&out)[U]) {
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
__device__ __forceinline__ void store_units(__nv_bfloat16* p, const float (&v)[U]) {
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

// Grid-wide exchange handshake: one arrival counter in L2; each CTA adds 1 (release) per published
// timestep and the TMA warp polls it (a per-CTA flag array polled lane-parallel was measured 2.5x
// slower: 9 000 vs 3 500 cycles from publish to release).
__device__ __forceinline__ uint32_t ld_relaxed_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Consumer side of the grid handshake.  The producers order their data stores before the counter
// increment (release: MEMBAR.GPU + RED), so once the counter value has been READ here, the data are
// already performed at L2.  The only consumer of that data is the TMA load issued below, which (a) is
// control-dependent on the value read, (b) follows a generic->async proxy fence and (c) reads L2, never
// this SM's L1.  A full gpu-scope acquire fence here (MEMBAR.ALL.GPU + CCTL.IVALL in SASS) therefore buys
// nothing the TMA needs and costs ~900 cycles per timestep (0.5 us of ~5), so it is optional
// (debug flag 16 re-enables it; tests/test_gpu_kernels.py::test_gru_grid_handshake_stress_bit_exact
// checks 10 M handshakes bit-exactly both ways).
// Measured: one polling iteration takes ~650 cycles inside the running kernel (an idle ld.relaxed.gpu round trip is
// ~280, probed from every CTA to 8 addresses: no die effect) and the CTAs leave the wait in two groups ~350 ns
// apart.  Neither replicating the counter over 8 L2 slices (every CTA adding to all, polling one), nor polling from
// all 4 issuing threads with staggered phases and a shared-memory claim (1/4 of the polling granularity; with one
// counter or one per thread index) made the step shorter: the release path (store ack + MEMBAR.GPU + RED), not the
// polling granularity, sets the ~1.2 us from the last publish to the first release.  Publishing per epilogue warp
// (4 release arrivals per CTA and step instead of a CTA barrier + one) is slower: 4.55 / 5.43 us per step.
// Watchdog: a protocol bug must end the launch with an error instead of hanging the GPU, but a spin COUNT can fire
// spuriously when the kernel is time-sliced (debugger, profiler replay, MPS), so the limit is wall-clock: 20 s without
// progress on one handshake (a timestep takes microseconds).
// `depth` polls are kept in flight (1, 2 or 4), spaced by a quarter of a poll's round trip: with one poll at a time the
// arrival of the last increment is noticed up to a full L2 round trip (~650 cycles under load) late.
__device__ __forceinline__ void spin_cycles(uint32_t n) {
  const long long t0 = clock64();
  while (clock64() - t0 < static_cast<long long>(n)) {
  }
}
__device__ __forceinline__ void grid_wait(const uint32_t* counter, uint32_t target, bool acquire_fence, int depth,
                                          uint32_t gap) {
  uint32_t spins = 0;
  unsigned long long t0 = 0;
  if (depth <= 1) {
    while (ld_relaxed_gpu(counter) < target) {
      if ((++spins & 0xFFFFu) == 0) {
        const unsigned long long now = globaltimer_ns();
        if (t0 == 0) t0 = now;
        else if (now - t0 > 20000000000ull) __trap();
      }
    }
  } else {
    // rotating window of asynchronous loads: a load only blocks when its value is consumed
    uint32_t v0, v1, v2 = 0, v3 = 0;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v0) : "l"(counter) : "memory");
    spin_cycles(gap);
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v1) : "l"(counter) : "memory");
    if (depth >= 4) {
      spin_cycles(gap);
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v2) : "l"(counter) : "memory");
      spin_cycles(gap);
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v3) : "l"(counter) : "memory");
    }
    for (;;) {
      if (v0 >= target) break;
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v0) : "l"(counter) : "memory");
      if (v1 >= target) break;
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v1) : "l"(counter) : "memory");
      if (depth >= 4) {
        if (v2 >= target) break;
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v2) : "l"(counter) : "memory");
        if (v3 >= target) break;
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v3) : "l"(counter) : "memory");
      }
      if ((++spins & 0x3FFFu) == 0) {
        const unsigned long long now = globaltimer_ns();
        if (t0 == 0) t0 = now;
        else if (now - t0 > 20000000000ull) __trap();
      }
    }
  }
  if (acquire_fence) asm volatile("fence.acq_rel.gpu;" ::: "memory");
}

// Producer side of the grid handshake: one arrival per CTA and published timestep, after a CTA barrier behind the data
// stores.  Default: RELAXED - the data stores and the increment travel to L2 concurrently and consumers validate what
// they read (see exchange()).  Tuning flag 16 = strict protocol: release increment (MEMBAR.ALL.GPU + RED) here and an
// acquire fence after the consumer's wait.
//
// Release fan-out (tuning flag 1 << 24, an experiment that LOST): the increment is an ATOM whose return value tells the
// CTA whether it was the LAST arriver of the round; that CTA alone then writes the round number into one flag line per
// CTA (sync[64 + 32 j], 128 B apart; 4 store instructions of one warp for 128 CTAs), and every CTA polls only its own
// line, so nobody reads the counter's L2 line.  Measured on one box, B=64 H=1024, us per timestep fwd / bwd: every CTA
// polling the counter 3.64 / 4.49, fan-out 4.29 / 5.00 - the 128 returning atomics serialise on the counter and the last
// arriver's return + flag stores + a second poll round trip cost more than the reader queue they remove.
// Called by the whole of epilogue warp 2 (converged); `round` = publishes of this CTA so far, this one included.
constexpr int GRU_FLAG_BASE = 64, GRU_FLAG_STRIDE = 32;            // in uint32 words of the sync buffer
// Per-group arrival counters (multi-group launches): group g of timestep t only needs group g's rows of h_{t-1}, so each
// group publishes to and waits on ITS OWN counter.  A CTA walks the groups of a timestep one after the other, so by the
// time it returns to group g for the next timestep, the other CTAs' publishes of group g are a whole group phase old:
// the grid handshake (~2 300 cycles) of one group hides behind the landing + MMAs + gate math of the others.
constexpr int GRU_GROUP_BASE = GRU_FLAG_BASE + GRU_FLAG_STRIDE * 128, GRU_GROUP_STRIDE = 32;
__device__ __forceinline__ uint32_t* group_counter(uint32_t* sync, int group_sync, int grp) {
  return group_sync ? sync + GRU_GROUP_BASE + GRU_GROUP_STRIDE * grp : sync;
}
__device__ __forceinline__ void publish(uint32_t* sync, int flags, uint32_t round, uint32_t n_ctas, int lane) {
  const bool strict = (flags & 16) != 0;
  if (!(flags & (1 << 24))) {                                    // default: plain counter, polled by everyone
    if (lane == 0) {
      if (strict) red_release_gpu_add(sync, 1u);
      else asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(sync), "r"(1u) : "memory");
    }
    return;
  }
  uint32_t old = 0;
  if (lane == 0) {
    if (strict) asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(sync), "r"(1u) : "memory");
    else asm volatile("atom.relaxed.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(sync), "r"(1u) : "memory");
  }
  old = __shfl_sync(0xffffffffu, old, 0);
  if (old + 1u == n_ctas * round) {                              // last arriver of this round: release everyone
    if (strict) asm volatile("fence.acq_rel.gpu;" ::: "memory");
    for (uint32_t j = static_cast<uint32_t>(lane); j < n_ctas; j += 32u)
      asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(sync + GRU_FLAG_BASE + GRU_FLAG_STRIDE * j), "r"(round) : "memory");
  }
}
constexpr uint32_t GRU_MAX_RETRIES = 256;

// CTA-wide OR over a named barrier (the epilogue warps only)
template <int ID, int THREADS>
__device__ __forceinline__ uint32_t bar_red_or(bool pred) {
  uint32_t out;
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "setp.ne.u32 q, %1, 0;\n"
      "barrier.cta.red.or.pred p, %2, %3, q;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(out)
      : "r"(static_cast<uint32_t>(pred)), "n"(ID), "n"(THREADS)
      : "memory");
  return out;
}
// remote arrival on a peer CTA's mbarrier / wait with cluster-scope acquire (the retry acknowledgement)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t remote_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  while (!ok) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if ((++spins & 0xFFFFFu) == 0 && spins > (1u << 30)) __trap();
  }
}

// Instrumented builds only (-DSRNN_DEBUG): timing experiments that produce WRONG results.  bit 0: do not wait on the
// grid counter; bit 2: skip the per-step global loads/stores of the epilogue.  A release build rejects both bits.
__device__ __forceinline__ bool grid_wait_skipped(int flags) {
#ifdef SRNN_DEBUG
  return (flags & 1) != 0;
#else
  (void)flags;
  return false;
#endif
}
__device__ __forceinline__ bool epilogue_io_skipped(int flags) {
#ifdef SRNN_DEBUG
  return (flags & 4) != 0;
#else
  (void)flags;
  return false;
#endif
}

// LSTM = false: GRU (gates r,z,n; 3H pre-activations).  LSTM = true: LSTM (gates i,f,g,o; 4H), an
// extension with no reference counterpart (BASELINE config 3; torch.nn.LSTM semantics, see oracle).
template <bool BWD, int C, bool LSTM, int U, int MW>
__global__ void __launch_bounds__(gru_threads(MW), 1)
gru_kernel(const __grid_constant__ CUtensorMap tma_w, const __grid_constant__ CUtensorMap tma_x, const GruParams p) {
  constexpr int UC = U * C;                          // units owned by the cluster
  constexpr int GATES = LSTM ? 4 : 3;
  constexpr int NG = BWD ? 1 : GATES;
  constexpr int NCOLS = NG * UC;                     // MMA N
  constexpr int NB = (NCOLS + 31) / 32;              // 32-column TMEM load batches per partial
  // one partial accumulator per issuing warp: partial w lives at columns [w*NCOLS, (w+1)*NCOLS); the
  // epilogue reads whole 32-column chunks, so the allocation covers the over-read of the last chunk
  constexpr int TMEM_NEED = (MW - 1) * NCOLS + NB * 32;
  constexpr uint32_t TMEM_COLS = TMEM_NEED <= 32 ? 32 : (TMEM_NEED <= 64 ? 64 : (TMEM_NEED <= 128 ? 128 :
                                 (TMEM_NEED <= 256 ? 256 : 512)));
  constexpr uint32_t IDESC = idesc_bf16(GRU_M, NCOLS, false, false);
  constexpr int WBLOCK = NCOLS * 128;                // resident weight bytes per K block

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int KBC = p.kbc;
  uint8_t* sw = smem;                                                   // [KBC][NCOLS rows][128 B]
  uint8_t* hbuf = sw + KBC * WBLOCK;                                    // [KBC][64 rows][128 B]
  float* part = reinterpret_cast<float*>(hbuf + KBC * GRU_SLOT);        // [NCOLS][64] fp32 partial tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(part + NCOLS * GRU_M);
  uint64_t* wfull = bars;
  uint64_t* acc_full = bars + 1;
  uint64_t* part_ready = bars + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
  uint64_t* retry_ack = bars + 4;                                       // C > 1: peers have consumed a rejected attempt's partials
  volatile uint32_t* ctl = reinterpret_cast<volatile uint32_t*>(bars + 5);   // [0] verdict (attempt << 1 | retry), [1] exit
  uint64_t* full = bars + 6;                                            // [GRU_MAX_KBC]: one per K block of the operand

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int H = p.hidden, T = p.steps, B = p.batch, EB = p.ext_batch;
  const uint32_t crank = C > 1 ? cluster_ctarank() : 0u;
  const int cluster_id = blockIdx.x / C;
  const int u0 = blockIdx.x * U;                     // first unit finalised by this CTA
  const int kb0 = static_cast<int>(crank) * KBC;     // first K block of this CTA's slice
  const uint32_t G = gridDim.x;
  // More than 64 batch rows per launch: the rows are walked in GROUPS of 64 inside every timestep (one MMA tile, one
  // landing of the operand per group) with ONE grid handshake per timestep; the weight slice stays resident for all
  // of them.  With several groups the fp32 recurrent state of a row lives in its global buffer between timesteps
  // (h_state / c_state forward, dh0 / dc0 as the running carries backward) instead of in registers.
  const int n_groups = p.groups;
  const int steps_end = BWD ? T + 1 : T;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_w);
    tma_prefetch_desc(&tma_x);
    mbar_init(wfull, 1);
    for (int kb = 0; kb < GRU_MAX_KBC; ++kb) mbar_init(full + kb, 1);
    mbar_init(acc_full, MW);
    mbar_init(part_ready, 1);
    mbar_init(retry_ack, C);
    ctl[0] = 0u;
    ctl[1] = 0u;
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (C > 1) cluster_sync_all();                     // every CTA's barriers exist before remote arrivals
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) {
    // resident weight slice: static data, so under programmatic dependent launch (single-step launches of
    // generation) it is fetched while the kernel that produces gi / h is still running
    mbar_expect_tx(wfull, static_cast<uint32_t>(KBC * WBLOCK));
    for (int kb = 0; kb < KBC; ++kb)
      for (int g = 0; g < NG; ++g)
        tma_load_2d(sw + kb * WBLOCK + g * UC * 128, &tma_w, wfull, (kb0 + kb) * 64,
                    (BWD ? 0 : g * H) + cluster_id * UC);   // gate g rows of W_hh (fwd)
  }
  pdl_launch_dependents();
  pdl_wait();                                        // no-ops for the cooperative multi-step launches of training

  if (warp == 0 || warp == 1 || warp >= 6) {
    // ------------------------------ TMA producer + MMA issuers ------------------------------
    // Thread 0 first waits for the grid, brings this CTA's K slice of the exchanged matrix, then joins the other
    // issuing threads (a CTA's load and its MMAs of one timestep are never concurrent, so the producer needs no
    // warp of its own).  The K steps of a block are dealt round-robin to the MW issuing threads, each with its
    // own partial accumulator (summed by the epilogue).
    if (lane == 0) {
      const int mw = warp < 2 ? warp : warp - 4;
      const uint32_t bytes = static_cast<uint32_t>(KBC) * static_cast<uint32_t>(p.box_rows) * 128u;
      mbar_wait(wfull, 0);
      const uint64_t a_base = smem_desc_sw128(smem_u32(hbuf), 16, 1024);
      const uint64_t b_base = smem_desc_sw128(smem_u32(sw), 16, 1024);
      const uint32_t a_hi = static_cast<uint32_t>(a_base >> 32), b_hi = static_cast<uint32_t>(b_base >> 32);
      const uint32_t d_tmem = tmem_base + mw * NCOLS;
      const uint32_t hslot16 = static_cast<uint32_t>(p.hslot) >> 4;
      const uint32_t a_lo0 = static_cast<uint32_t>(a_base), b_lo0 = static_cast<uint32_t>(b_base);
      const uint32_t blk_bytes = static_cast<uint32_t>(p.box_rows) * 128u;
      const bool one_box = p.one_box != 0;
      const bool strict = (p.flags & 16) != 0;
      // tuning (experiments): bits 12-13 polls in flight (0 -> 1, 1 -> 2, 2 -> 4), bits 14-15 their spacing
      // ((n + 1) * 64 cycles), bits 16-19 cycles/32 to hold back the TMA read after the wait (fewer rejected attempts)
      const int poll_depth = 1 << ((p.flags >> 12) & 3);
      const uint32_t poll_gap = (((p.flags >> 14) & 3) + 1) * 64u;
      const uint32_t hold = ((p.flags >> 16) & 15) * 32u;
      // bits 20-23: cycles before the FIRST poll of a wait (0: default 256, n: (n-1)*128).  This thread reaches the wait ~1 000 cycles before the
      // last CTA can have published (own gate math + publish skew); polling during that time only queues reads on the
      // counter's L2 line in front of the other CTAs' increments
      // (measured, B=64 H=1024: 0 cycles 3.68 / 4.33 us per step fwd / bwd, 256 cycles 3.54 / 4.26, 512 and more: no gain)
      // (aligning the TMA issue of all CTAs to a common global-timer boundary of 128 / 256 / 512 ns - on the theory that L2
      // merges same-line requests of different SMs only when they arrive together - measured 3.68 / 3.77 / 3.93 us per
      // forward timestep against 3.64: the landing stays ~1 900 cycles, the wait is pure cost)
      const uint32_t pp = (p.flags >> 20) & 15;
      const uint32_t pre_poll = pp == 0 ? 256u : (pp - 1) * 128u;
      // An ATTEMPT = landing of the operand + the MMAs + one commit.  The epilogue validates every attempt (see
      // exchange() below) and posts a verdict; a rejected attempt is repeated for the same timestep.  Thread 0 drives:
      // it knows the round, waits for the grid, issues the loads and reads the verdicts; the other issuing threads just
      // follow the landing barriers until thread 0 raises the exit flag.
      uint32_t phase = 0, attempt = 0;
      int s = BWD ? 1 : 0;                           // backward: the last timestep (round 0) has no recurrent input
      int grp = 0;                                   // group of 64 batch rows inside timestep s
      bool fresh = true;                             // first attempt of (s, grp)
      // Tuning flag 1 << 25, SPECULATIVE landing: the first attempt of a timestep does not wait for the arrival counter at
      // all - it issues the TMA read `pre_poll` cycles after this CTA's own publish and lets the validation decide; only a
      // rejected attempt falls back to the counter.  (The counter's round trip - RED to a hot L2 line + a poll - is longer
      // than the time the data stores themselves need to reach L2.)
      const bool speculate = !strict && (p.flags & (1 << 25)) != 0;
      bool waited = false;                           // this timestep's counter wait has been done
      for (;;) {
        if (mw == 0) {
          // one wait per timestep (before its first group), or - with per-group counters - one per group
          const bool head = fresh && (grp == 0 || p.group_sync);
          if (head) waited = false;
          const bool first = head && s > 0;
          if (s > 0 && !waited && !grid_wait_skipped(p.flags) && (speculate ? !fresh : first)) {
            if (pre_poll && first && grp == 0) spin_cycles(pre_poll);
            if (!(p.flags & (1 << 24)))
              grid_wait(group_counter(p.sync, p.group_sync, grp), G * static_cast<uint32_t>(s), strict, poll_depth, poll_gap);
            else                                     // experiment: own flag line, written by the last arriver of round s
              grid_wait(p.sync + GRU_FLAG_BASE + GRU_FLAG_STRIDE * blockIdx.x, static_cast<uint32_t>(s), strict, poll_depth,
                        poll_gap);
            if (hold) spin_cycles(hold);
            waited = true;
            GRU_TS(0, s);
          } else if (speculate && first) {
            if (pre_poll) spin_cycles(pre_poll);
            GRU_TS(0, s);
          }
          asm volatile("fence.proxy.async.global;" ::: "memory");     // generic-proxy writes -> TMA reads (free: measured)
          const int slot = BWD ? (T - s) : s;                     // time slot of the exchange buffer
          const int row0 = grp * p.group_rows;                    // first batch row of this group
          if (one_box) {
            mbar_expect_tx(full, bytes);
            tma_load_4d(hbuf, &tma_x, full, 0, row0, kb0, slot);
          } else {
            for (int kb = 0; kb < KBC; ++kb) {       // tuning flag 32: one box and one barrier per K block
              mbar_expect_tx(full + kb, blk_bytes);
              tma_load_3d(hbuf + kb * GRU_SLOT, &tma_x, full + kb, (kb0 + kb) * 64, row0, slot);
            }
          }
          GRU_TS(1, s);
          mbar_wait(full, phase);
        } else {
          bool leave = false;
          while (!mbar_try_wait(full, phase)) {
            if (ctl[1] != 0u) {
              leave = true;
              break;
            }
          }
          if (leave) break;
        }
        if (mw == 0) GRU_TS(2, s);
        tc_fence_after();
        // K steps (16 columns = 32 bytes inside a 64-column block) are dealt round-robin: issuer mw takes steps
        // mw, mw + MW, ...  Only the 14-bit start-address field (low descriptor word) changes.  (Keep this loop lean:
        // one thread issues it, ~5 cycles per dependent SASS instruction - 45 instructions per MMA cost 2.5x the step.)
        if (one_box) {
#pragma unroll 4
          for (int ks = mw; ks < KBC * 4; ks += MW) {
            const uint32_t kb = static_cast<uint32_t>(ks) >> 2, j = static_cast<uint32_t>(ks) & 3u;
            const uint64_t ad = (static_cast<uint64_t>(a_hi) << 32) | (a_lo0 + kb * hslot16 + j * 2u);
            const uint64_t bd = (static_cast<uint64_t>(b_hi) << 32) | (b_lo0 + kb * (WBLOCK >> 4) + j * 2u);
            umma_bf16(d_tmem, ad, bd, IDESC, ks >= MW ? 1u : 0u);
          }
        } else {
          for (int ks = mw; ks < KBC * 4; ks += MW) {
            const uint32_t kb = static_cast<uint32_t>(ks) >> 2, j = static_cast<uint32_t>(ks) & 3u;
            if (kb != 0u && j < static_cast<uint32_t>(MW)) mbar_wait(full + kb, phase);
            const uint64_t ad = (static_cast<uint64_t>(a_hi) << 32) | (a_lo0 + kb * hslot16 + j * 2u);
            const uint64_t bd = (static_cast<uint64_t>(b_hi) << 32) | (b_lo0 + kb * (WBLOCK >> 4) + j * 2u);
            umma_bf16(d_tmem, ad, bd, IDESC, ks >= MW ? 1u : 0u);
          }
        }
        umma_commit(acc_full);
        phase ^= 1;
        ++attempt;
        if (mw == 0) {
          GRU_TS(3, s);
          uint32_t v, spins = 0;
          unsigned long long w0 = 0;
          while (((v = ctl[0]) >> 1) != attempt) {   // the epilogue's verdict on this attempt
            if ((++spins & 0xFFFFFu) == 0) {
              const unsigned long long now = globaltimer_ns();
              if (w0 == 0) w0 = now;
              else if (now - w0 > 20000000000ull) __trap();
            }
          }
          tc_fence_after();
          if (v & 1u) {
            fresh = false;                           // rejected (the operand was not complete yet): same round again
            continue;
          }
          fresh = true;
          if (++grp == n_groups) {
            grp = 0;
            if (++s == steps_end) {
              ctl[1] = 1u;
              break;
            }
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------ epilogue ------------------------------
    const int q = warp & 3;                          // TMEM lane quadrant of this warp
    const int row = q * 16 + lane;                   // M=64: rows 16q..16q+15 live in lanes 32q..32q+15
    const bool lane_ok = lane < 16;
    // `row` is the row inside the 64-row MMA tile, `grow` the batch row it stands for (they differ in multi-group mode)
    int grow = row;
    bool row_ok = lane_ok && grow < B;
    bool io = row_ok && !epilogue_io_skipped(p.flags);
    const bool regs = n_groups == 1;                 // the recurrent state stays in registers across timesteps
    // (A "quiet window" - the epilogue warps holding their off-critical-path stores and the next step's input loads back for
    // 512..3 072 cycles after the publish so that this SM's memory pipeline stays free for the polls - measured 3.76-3.99 us
    // per forward timestep against 3.60 and no change backward: the spinning warps take issue slots from the issuing thread.)
    auto set_group = [&](int g) {
      grow = g * p.group_rows + row;
      row_ok = lane_ok && row < p.group_rows && grow < B;
      io = row_ok && !epilogue_io_skipped(p.flags);
    };
    const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t part_addr = smem_u32(part);
    const uint32_t ready_addr = smem_u32(part_ready);
    uint32_t acc_phase = 0, part_phase = 0;

    // Cluster reduction, push model: every rank sends, for each destination rank, the 8 columns of
    // that rank's units straight into the destination's shared memory with st.async (layout
    // recv[src rank][gate][row][8 units]); the bytes complete on the destination's mbarrier, so there
    // is no staging buffer, no cluster-scope fence and no remote-load latency on the critical path.
    // WAR safety: a peer pushes step t+1 only after the grid barrier of step t, i.e. after this CTA
    // has consumed step t.
    constexpr uint32_t RECV_BYTES = C * NG * GRU_M * U * 4;
    uint32_t attempt = 0, ack_phase = 0;
    // Validation of the exchanged operand.  Producers publish with plain stores followed by a RELAXED counter increment
    // (no MEMBAR.GPU: it cost ~1 us of every ~4 us step), so the counter may become visible before the data.  The
    // exchange buffer is pre-filled with bf16 NaN (0xFFFF) by the host wrapper and every element is written exactly once
    // per launch; an element that has not arrived yet therefore turns ALL accumulators of its batch row into NaN
    // (tensor cores propagate NaN; checked by scripts/nan_probe.py).  One accumulator per row is tested, the CTA votes,
    // and a rejected attempt is repeated (reload + MMAs) for the same timestep.  After GRU_MAX_RETRIES rejections the
    // NaN is accepted as genuine state (diverged training) so the launch always terminates.  With C > 1 all CTAs of
    // a cluster see the same rows as NaN (each sums the partials of all C K-slices), so they decide alike.
    auto exchange = [&](float (&out)[NG * U], int dbg_step) {
      uint32_t retries = 0;
      for (;;) {
        if constexpr (C == 1) {                          // no peers: the accumulator columns are the result
          mbar_wait(acc_full, acc_phase);
          acc_phase ^= 1;
          tc_fence_after();
          static_assert(C != 1 || NCOLS <= 32, "C == 1: one 32-column chunk per partial");
#pragma unroll
          for (int i = 0; i < NG * U; ++i) out[i] = 0.f;
          if constexpr (NG * U == 24 && MW == 4) {
            // GRU forward: the 24 columns of all four partial tiles are requested at once (16 + 8 columns each, 96
            // registers) and waited for ONCE - one TMEM round trip per timestep instead of two
            uint32_t va[MW][16], vb[MW][8];
#pragma unroll
            for (int pw = 0; pw < MW; ++pw) {
              tmem_ld16(t_addr + pw * NCOLS, va[pw]);
              tmem_ld8(t_addr + pw * NCOLS + 16, vb[pw]);
            }
            tmem_ld_wait();
#pragma unroll
            for (int pw = 0; pw < MW; ++pw) {
#pragma unroll
              for (int i = 0; i < 16; ++i) out[i] += __uint_as_float(va[pw][i]);
#pragma unroll
              for (int i = 0; i < 8; ++i) out[16 + i] += __uint_as_float(vb[pw][i]);
            }
          } else {
#pragma unroll
            for (int pw = 0; pw < MW; pw += 2) {           // two partial tiles in flight per wait
              if (pw >= KBC * 4) break;                    // fewer K steps than issuers: those partials were never written
              uint32_t v[2][32];
              tmem_ld32(t_addr + pw * NCOLS, v[0]);
              tmem_ld32(t_addr + (pw + 1) * NCOLS, v[1]);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < NG * U; ++i) out[i] += __uint_as_float(v[0][i]) + __uint_as_float(v[1][i]);
            }
          }
          tc_fence_before();
        } else {
          if (warp == 2 && lane == 0) mbar_expect_tx(part_ready, RECV_BYTES);
          mbar_wait(acc_full, acc_phase);
          acc_phase ^= 1;
          if (warp == 2 && lane == 0) GRU_TS(4, dbg_step);
          tc_fence_after();
#pragma unroll
          for (int bi = 0; bi < NB; ++bi) {                // 32 result columns per batch, ONE wait per batch
            uint32_t v[MW][32];
#pragma unroll
            for (int pw = 0; pw < MW; ++pw) tmem_ld32(t_addr + pw * NCOLS + bi * 32, v[pw]);
            tmem_ld_wait();
            if (lane_ok) {
#pragma unroll
              for (int e = 0; e < 4; ++e) {                 // groups of 8 columns = one (gate, dst rank) pair
                const int col = bi * 32 + e * 8;
                if (col < NCOLS) {
                  float f[8];
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    f[i] = 0.f;
#pragma unroll
                    for (int pw = 0; pw < MW; ++pw) f[i] += __uint_as_float(v[pw][e * 8 + i]);
                  }
                  const int g = col / UC, dst = (col % UC) / U, sub = (col % U) / 8;
                  const uint32_t off = static_cast<uint32_t>((((crank * NG + g) * GRU_M + row) * U + sub * 8) * 4);
                  const uint32_t ra = mapa(part_addr + off, static_cast<uint32_t>(dst));
                  const uint32_t rb = mapa(ready_addr, static_cast<uint32_t>(dst));
                  st_async_v4(ra, f[0], f[1], f[2], f[3], rb);
                  st_async_v4(ra + 16, f[4], f[5], f[6], f[7], rb);
                }
              }
            }
          }
          tc_fence_before();
          if (warp == 2 && lane == 0) GRU_TS(7, dbg_step);
          mbar_wait(part_ready, part_phase);
          part_phase ^= 1;
          if (warp == 2 && lane == 0) GRU_TS(5, dbg_step);
#pragma unroll
          for (int i = 0; i < NG * U; ++i) out[i] = 0.f;
          if (lane_ok) {
#pragma unroll
            for (int r = 0; r < C; ++r)
#pragma unroll
              for (int g = 0; g < NG; ++g) {
                const float4* rp = reinterpret_cast<const float4*>(part + ((r * NG + g) * GRU_M + row) * U);
#pragma unroll
                for (int i4 = 0; i4 < U / 4; ++i4) {
                  const float4 a = rp[i4];
                  out[g * U + i4 * 4 + 0] += a.x; out[g * U + i4 * 4 + 1] += a.y;
                  out[g * U + i4 * 4 + 2] += a.z; out[g * U + i4 * 4 + 3] += a.w;
                }
              }
          }
        }
        ++attempt;
        const bool bad = row_ok && (out[0] != out[0]);
        uint32_t reject = bar_red_or<2, 128>(bad);
        if (reject && retries >= GRU_MAX_RETRIES) reject = 0u;
        if (warp == 2 && lane == 0) {
          ctl[0] = (attempt << 1) | reject;
          if (reject) atomicAdd(p.sync + 32, 1u);        // statistics: rejected attempts of this launch (sync[32])
        }
        if (!reject) return;
        ++retries;
        if constexpr (C > 1) {
          // nobody may push the next attempt's partials into a peer that is still summing this attempt's
          if (warp == 2 && lane == 0) {
            const uint32_t ack_addr = smem_u32(retry_ack);
#pragma unroll
            for (int r = 0; r < C; ++r) mbar_arrive_cluster(mapa(ack_addr, static_cast<uint32_t>(r)));
          }
          mbar_wait_cluster(retry_ack, ack_phase);
          ack_phase ^= 1;
        }
      }
    };

    if constexpr (!BWD && LSTM) {
      // ---------------- LSTM forward: c' = f c + i g,  h' = o tanh(c') ----------------
      float h[U], c[U], bi[U], bf[U], bg[U], bo[U];
#pragma unroll
      for (int i = 0; i < U; ++i) {
        h[i] = (regs && row_ok) ? p.h_state[static_cast<long long>(grow) * H + u0 + i] : 0.f;
        c[i] = (regs && row_ok) ? p.c_state[static_cast<long long>(grow) * H + u0 + i] : 0.f;
        bi[i] = p.b_hh[u0 + i];
        bf[i] = p.b_hh[H + u0 + i];
        bg[i] = p.b_hh[2 * H + u0 + i];
        bo[i] = p.b_hh[3 * H + u0 + i];
      }
      for (int t = 0; t < T; ++t) {
        for (int grp = 0; grp < n_groups; ++grp) {
          if (!regs) set_group(grp);
          const long long rt = static_cast<long long>(grow) * T + t;
          float xi[U], xf[U], xg[U], xo[U];
#pragma unroll
          for (int i = 0; i < U; ++i) xi[i] = xf[i] = xg[i] = xo[i] = 0.f;
          if (io) {
            const __nv_bfloat16* gp = p.gi + rt * 4 * H + u0;
            load_units<U>(gp, xi);
            load_units<U>(gp + H, xf);
            load_units<U>(gp + 2 * H, xg);
            load_units<U>(gp + 3 * H, xo);
          }
          if (!regs) {
#pragma unroll
            for (int i = 0; i < U; ++i) {
              h[i] = row_ok ? p.h_state[static_cast<long long>(grow) * H + u0 + i] : 0.f;
              c[i] = row_ok ? p.c_state[static_cast<long long>(grow) * H + u0 + i] : 0.f;
            }
          }
          float acc[4 * U];
          exchange(acc, t);
          float gi_[U], gf_[U], gg_[U], go_[U];
#pragma unroll
          for (int i = 0; i < U; ++i) {
            gi_[i] = sigmoid_fast(xi[i] + acc[i] + bi[i]);
            gf_[i] = sigmoid_fast(xf[i] + acc[U + i] + bf[i]);
            gg_[i] = tanh_fast(xg[i] + acc[2 * U + i] + bg[i]);
            go_[i] = sigmoid_fast(xo[i] + acc[3 * U + i] + bo[i]);
            c[i] = gf_[i] * c[i] + gi_[i] * gg_[i];
            h[i] = go_[i] * tanh_fast(c[i]);
          }
          if (io) store_units<U>(p.h_ext + (static_cast<long long>(t + 1) * EB + grow) * H + u0, h);
          if (!regs && row_ok) {
#pragma unroll
            for (int i = 0; i < U; ++i) {
              p.h_state[static_cast<long long>(grow) * H + u0 + i] = h[i];
              p.c_state[static_cast<long long>(grow) * H + u0 + i] = c[i];
            }
          }
          if (p.group_sync || grp == n_groups - 1) {                   // every group's h_t is stored: one arrival per CTA and timestep
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (warp == 2 && T > 1) publish(group_counter(p.sync, p.group_sync, grp), p.flags, static_cast<uint32_t>(t + 1), G, lane);   // T == 1: nobody waits
          }
          if (p.hall && io) store_units<U>(p.hall + rt * H + u0, h);
          if (p.gates && io) {                           // saved for backward: i, f, g, o, c_t  (5H per row)
            __nv_bfloat16* sp = p.gates + rt * 5 * H + u0;
            store_units<U>(sp, gi_);
            store_units<U>(sp + H, gf_);
            store_units<U>(sp + 2 * H, gg_);
            store_units<U>(sp + 3 * H, go_);
            store_units<U>(sp + 4 * H, c);
          }
        }
      }
      if (regs && row_ok) {
#pragma unroll
        for (int i = 0; i < U; ++i) {
          p.h_state[static_cast<long long>(grow) * H + u0 + i] = h[i];
          p.c_state[static_cast<long long>(grow) * H + u0 + i] = c[i];
        }
      }
    } else if constexpr (BWD && LSTM) {
      // ---------------- LSTM backward ----------------
      float carry_c[U];
      float sb[4 * U];                                 // bias gradients: running sums of the 4 gate gradients of this thread's rows
#pragma unroll
      for (int i = 0; i < U; ++i) carry_c[i] = 0.f;
#pragma unroll
      for (int i = 0; i < 4 * U; ++i) sb[i] = 0.f;
      for (int s = 0; s <= T; ++s) {
        const int t = T - 1 - s;
        for (int grp = 0; grp < n_groups; ++grp) {
          if (!regs) set_group(grp);
          const long long rt = static_cast<long long>(grow) * T + t;
          float dh[U], gi_[U], gf_[U], gg_[U], go_[U], ct[U], cp[U];
#pragma unroll
          for (int i = 0; i < U; ++i) dh[i] = gi_[i] = gf_[i] = gg_[i] = go_[i] = ct[i] = cp[i] = 0.f;
          if (io && t >= 0) {
            load_units<U>(p.dh_out + rt * H + u0, dh);
            const __nv_bfloat16* sp = p.gates + rt * 5 * H + u0;
            load_units<U>(sp, gi_);
            load_units<U>(sp + H, gf_);
            load_units<U>(sp + 2 * H, gg_);
            load_units<U>(sp + 3 * H, go_);
            load_units<U>(sp + 4 * H, ct);
            if (t > 0) {
              load_units<U>(p.gates + (rt - 1) * 5 * H + 4 * H + u0, cp);
            } else {
#pragma unroll
              for (int i = 0; i < U; ++i) cp[i] = p.c_init[static_cast<long long>(grow) * H + u0 + i];
            }
          }
          if (!regs) {                                   // the running cell-state carry of this row lives in dc0
#pragma unroll
            for (int i = 0; i < U; ++i)
              carry_c[i] = (s > 0 && row_ok) ? p.dc0[static_cast<long long>(grow) * H + u0 + i] : 0.f;
          }
          float d[U];
#pragma unroll
          for (int i = 0; i < U; ++i) d[i] = 0.f;
          if (s > 0) exchange(d, s);
          if (t < 0) {
            if (row_ok) {
#pragma unroll
              for (int i = 0; i < U; ++i) {
                p.dh0[static_cast<long long>(grow) * H + u0 + i] = d[i];
                p.dc0[static_cast<long long>(grow) * H + u0 + i] = carry_c[i];
              }
            }
            continue;
          }
          float pi[U], pf[U], pg[U], po[U];
#pragma unroll
          for (int i = 0; i < U; ++i) {
            const float dht = dh[i] + d[i];
            const float tc = tanh_fast(ct[i]);
            const float dc = dht * go_[i] * (1.f - tc * tc) + carry_c[i];
            carry_c[i] = dc * gf_[i];
            pi[i] = dc * gg_[i] * gi_[i] * (1.f - gi_[i]);
            pf[i] = dc * cp[i] * gf_[i] * (1.f - gf_[i]);
            pg[i] = dc * gi_[i] * (1.f - gg_[i] * gg_[i]);
            po[i] = dht * tc * go_[i] * (1.f - go_[i]);
          }
          if (io) {
            __nv_bfloat16* ghp = p.dgh + (static_cast<long long>(t) * EB + grow) * 4 * H + u0;   // exchanged
            store_units<U>(ghp, pi);
            store_units<U>(ghp + H, pf);
            store_units<U>(ghp + 2 * H, pg);
            store_units<U>(ghp + 3 * H, po);
          }
          if (!regs && row_ok) {
#pragma unroll
            for (int i = 0; i < U; ++i) p.dc0[static_cast<long long>(grow) * H + u0 + i] = carry_c[i];
          }
          if (p.group_sync || grp == n_groups - 1) {
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (warp == 2) publish(group_counter(p.sync, p.group_sync, grp), p.flags, static_cast<uint32_t>(s + 1), G, lane);
          }
          if (io) {                                               // same values, batch-major, for the GEMMs
            __nv_bfloat16* gip = p.dgi + rt * 4 * H + u0;
            store_units<U>(gip, pi);
            store_units<U>(gip + H, pf);
            store_units<U>(gip + 2 * H, pg);
            store_units<U>(gip + 3 * H, po);
          }
          if (row_ok) {
#pragma unroll
            for (int i = 0; i < U; ++i) {
              sb[i] += pi[i]; sb[U + i] += pf[i]; sb[2 * U + i] += pg[i]; sb[3 * U + i] += po[i];
            }
          }
        }
      }
      if (p.db_ih || p.db_hh) {                                   // both biases see the same gate gradients
        float* sred = reinterpret_cast<float*>(hbuf);
        if (lane_ok) {
#pragma unroll
          for (int i = 0; i < 4 * U; ++i) sred[i * GRU_M + row] = sb[i];     // (0 for rows that never held a slot)
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int v = threadIdx.x - 64;
        if (v < 4 * U) {
          float acc = 0.f;
          for (int r = 0; r < GRU_M; ++r) acc += sred[v * GRU_M + r];
          const int g = v / U, i = v % U;
          if (p.db_ih) p.db_ih[g * H + u0 + i] += acc;
          if (p.db_hh) p.db_hh[g * H + u0 + i] += acc;
        }
      }
    } else if constexpr (!BWD) {
      float h[U], bhr[U], bhz[U], bhn[U];
#pragma unroll
      for (int i = 0; i < U; ++i) {
        h[i] = (regs && row_ok) ? p.h_state[static_cast<long long>(grow) * H + u0 + i] : 0.f;
        bhr[i] = p.b_hh[u0 + i];
        bhz[i] = p.b_hh[H + u0 + i];
        bhn[i] = p.b_hh[2 * H + u0 + i];
      }
      for (int t = 0; t < T; ++t) {
        for (int grp = 0; grp < n_groups; ++grp) {
          if (!regs) set_group(grp);
          const long long rt = static_cast<long long>(grow) * T + t;
          float gr[U], gz[U], gn[U];
#pragma unroll
          for (int i = 0; i < U; ++i) gr[i] = gz[i] = gn[i] = 0.f;
          if (io) {
            const __nv_bfloat16* gp = p.gi + rt * 3 * H + u0;
            load_units<U>(gp, gr);
            load_units<U>(gp + H, gz);
            load_units<U>(gp + 2 * H, gn);
          }
          if (!regs) {
#pragma unroll
            for (int i = 0; i < U; ++i) h[i] = row_ok ? p.h_state[static_cast<long long>(grow) * H + u0 + i] : 0.f;
          }
          float acc[3 * U];
          exchange(acc, t);
          float r[U], z[U], n[U], hn[U];
#pragma unroll
          for (int i = 0; i < U; ++i) {
            r[i] = sigmoid_fast(gr[i] + acc[i] + bhr[i]);
            z[i] = sigmoid_fast(gz[i] + acc[U + i] + bhz[i]);
            hn[i] = acc[2 * U + i] + bhn[i];
            n[i] = tanh_fast(gn[i] + r[i] * hn[i]);
            h[i] = (1.f - z[i]) * n[i] + z[i] * h[i];
          }
          if (io) store_units<U>(p.h_ext + (static_cast<long long>(t + 1) * EB + grow) * H + u0, h);
          if (!regs && row_ok) {
#pragma unroll
            for (int i = 0; i < U; ++i) p.h_state[static_cast<long long>(grow) * H + u0 + i] = h[i];
          }
          if (p.group_sync || grp == n_groups - 1) {
            // publish h_t: all epilogue threads' stores (of every group) -> one arrival per CTA
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (warp == 2) {
              if (lane == 0) GRU_TS(6, t);
              if (T > 1) publish(group_counter(p.sync, p.group_sync, grp), p.flags, static_cast<uint32_t>(t + 1), G, lane);   // T == 1: nobody waits
            }
          }
          // the batch-major copy of h_t (GEMM operand) and the saved gates are off the critical path
          if (p.hall && io) store_units<U>(p.hall + rt * H + u0, h);
          if (p.gates && io) {
            __nv_bfloat16* sp = p.gates + rt * 4 * H + u0;
            store_units<U>(sp, r);
            store_units<U>(sp + H, z);
            store_units<U>(sp + 2 * H, n);
            store_units<U>(sp + 3 * H, hn);
          }
        }
      }
      if (regs && row_ok) {
#pragma unroll
        for (int i = 0; i < U; ++i) p.h_state[static_cast<long long>(grow) * H + u0 + i] = h[i];
      }
    } else {
      float carry[U];
      float sb[4 * U];                                 // bias gradients: running sums of gr, gz, gn, ghn of this thread's rows
#pragma unroll
      for (int i = 0; i < U; ++i) carry[i] = 0.f;
#pragma unroll
      for (int i = 0; i < 4 * U; ++i) sb[i] = 0.f;
      for (int s = 0; s <= T; ++s) {
        const int t = T - 1 - s;
        for (int grp = 0; grp < n_groups; ++grp) {
          if (!regs) set_group(grp);
          const long long rt = static_cast<long long>(grow) * T + t;
          float dh[U], r[U], z[U], n[U], hn[U], hp[U];
#pragma unroll
          for (int i = 0; i < U; ++i) dh[i] = r[i] = z[i] = n[i] = hn[i] = hp[i] = 0.f;
          if (io && t >= 0) {
            load_units<U>(p.dh_out + rt * H + u0, dh);
            const __nv_bfloat16* sp = p.gates + rt * 4 * H + u0;
            load_units<U>(sp, r);
            load_units<U>(sp + H, z);
            load_units<U>(sp + 2 * H, n);
            load_units<U>(sp + 3 * H, hn);
            load_units<U>(p.h_ext + (static_cast<long long>(t) * EB + grow) * H + u0, hp);
          }
          if (!regs) {                                   // the running carry z * dL/dh of this row lives in dh0
#pragma unroll
            for (int i = 0; i < U; ++i)
              carry[i] = (s > 0 && row_ok) ? p.dh0[static_cast<long long>(grow) * H + u0 + i] : 0.f;
          }
          float d[U];
#pragma unroll
          for (int i = 0; i < U; ++i) d[i] = 0.f;
          if (s > 0) exchange(d, s);
          if (t < 0) {
            if (row_ok) {
#pragma unroll
              for (int i = 0; i < U; ++i) p.dh0[static_cast<long long>(grow) * H + u0 + i] = carry[i] + d[i];
            }
            continue;
          }
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
          if (io) {
            __nv_bfloat16* ghp = p.dgh + (static_cast<long long>(t) * EB + grow) * 3 * H + u0;   // exchanged
            store_units<U>(ghp, gr);
            store_units<U>(ghp + H, gz);
            store_units<U>(ghp + 2 * H, ghn);
          }
          if (!regs && row_ok) {
#pragma unroll
            for (int i = 0; i < U; ++i) p.dh0[static_cast<long long>(grow) * H + u0 + i] = carry[i];
          }
          if (p.group_sync || grp == n_groups - 1) {
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (warp == 2) {
              if (lane == 0) GRU_TS(6, s);
              publish(group_counter(p.sync, p.group_sync, grp), p.flags, static_cast<uint32_t>(s + 1), G, lane);
            }
          }
          if (io) {                                               // dgi is only read after the kernel
            __nv_bfloat16* gip = p.dgi + rt * 3 * H + u0;
            store_units<U>(gip, gr);
            store_units<U>(gip + H, gz);
            store_units<U>(gip + 2 * H, gn);
          }
          if (row_ok) {                                           // (rows beyond the batch hold garbage)
#pragma unroll
            for (int i = 0; i < U; ++i) {
              sb[i] += gr[i]; sb[U + i] += gz[i]; sb[2 * U + i] += gn[i]; sb[3 * U + i] += ghn[i];
            }
          }
        }
      }
      // bias gradients: sum the per-row sums over the batch rows through shared memory (the operand buffer is free now)
      if (p.db_ih || p.db_hh) {
        float* sred = reinterpret_cast<float*>(hbuf);            // [4U][64 rows]
        if (lane_ok) {
#pragma unroll
          for (int i = 0; i < 4 * U; ++i) sred[i * GRU_M + row] = sb[i];     // (0 for rows that never held a slot)
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int v = threadIdx.x - 64;                          // epilogue threads 0..127
        if (v < 4 * U) {
          float acc = 0.f;
          for (int r = 0; r < GRU_M; ++r) acc += sred[v * GRU_M + r];
          const int g = v / U, i = v % U;                        // 0 r, 1 z, 2 n (input side), 3 n (hidden side)
          if (g < 3 && p.db_ih) p.db_ih[g * H + u0 + i] += acc;
          if (g != 2 && p.db_hh) p.db_hh[(g == 3 ? 2 : g) * H + u0 + i] += acc;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (C > 1) cluster_sync_all();                     // nobody exits while a peer may still read its smem
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <bool BWD, int C, bool LSTM, int U>
static size_t gru_smem_bytes(int kbc) {
  const int ncols = (BWD ? 1 : (LSTM ? 4 : 3)) * U * C;
  return static_cast<size_t>(kbc) * ncols * 128 + static_cast<size_t>(kbc) * GRU_SLOT +
         static_cast<size_t>(ncols) * GRU_M * 4 + (8 + GRU_MAX_KBC) * 8 + 1024;
}

// MMA issuing warps: 4 when the 4 partial accumulators fit the 512 TMEM columns, else 2
template <bool BWD, int C, bool LSTM, int U>
constexpr int gru_mw() {
  constexpr int ncols = (BWD ? 1 : (LSTM ? 4 : 3)) * U * C;
  if (U != 8) return 2;                                // 16 units per thread: 4 partial tiles would spill
  // (8 issuers at C=1 were measured no faster than 4: 64 MMAs in ~1550 cycles either way, i.e. the ~24-cycle
  // occupancy of the tensor pipe per M=64 instruction is the bound, no longer the issue rate)
  return 3 * ncols + ((ncols + 31) / 32) * 32 <= 512 ? 4 : 2;
}

// Can H/8 CTAs in clusters of C all be resident at once (they spin on one another)?
template <bool BWD, int C, bool LSTM, int U>
static bool gru_fits(int H, int kbc) {
  constexpr int MW = gru_mw<BWD, C, LSTM, U>();
  constexpr int GRU_THREADS = gru_threads(MW);
  auto kern = gru_kernel<BWD, C, LSTM, U, MW>;
  const size_t smem = gru_smem_bytes<BWD, C, LSTM, U>(kbc);
  if (smem > 227 * 1024) return false;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  const int ctas = H / U;
  if (C == 1) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, GRU_THREADS, smem) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    return per_sm * sm_count() >= ctas;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(GRU_THREADS);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int clusters = 0;
  if (cudaOccupancyMaxActiveClusters(&clusters, kern, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return clusters * C >= ctas;
}

template <bool BWD, int C, bool LSTM, int U, int MW>
static int launch_gru_mw(const srnn_gru_args* a, int kbc, cudaStream_t stream) {
  constexpr int GATES = LSTM ? 4 : 3;
  constexpr int GRU_THREADS = gru_threads(MW);
  const int H = a->hidden, T = a->steps, B = a->batch;
  const int K = BWD ? GATES * H : H;
  const int ctas = H / U;
  const size_t smem = gru_smem_bytes<BWD, C, LSTM, U>(kbc);

  CUtensorMap tw, tx;
  {
    // forward: W_hh [3H, H] rows = gate rows; backward: W_hh^T [H, 3H] rows = units
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)(BWD ? H : GATES * H)};
    const uint64_t strides[1] = {(uint64_t)K * 2};
    const uint32_t box[2] = {64, (uint32_t)(U * C)};
    int rc = make_tmap_bf16(&tw, a->w_hh, 2, dims, strides, box, true);
    if (rc) return rc;
  }
  // exchange buffer, TIME-major: forward h_ext [T+1, EB, H]; backward dgh [T, EB, 3H].  When K is a whole
  // number of 64-column blocks and the batch a whole number of 8-row swizzle atoms, the buffer is viewed as
  // [slot][K block][row][64] (strides not monotonic: legal for a tiled map) so that ONE box brings the CTA's
  // whole [batch, K slice] operand, K block by K block, in the layout the MMA reads (8 TMA issues -> 1).
  // Landing of the per-step operand: ONE 4-D TMA box (one issue, ~240 cycles) when the shape allows it.  Tuning flag 32
  // lands it as one box + one barrier per K block instead so that the MMAs of block i could start while later blocks
  // are in flight - measured slower (16 issues cost 1 500 cycles and the FIRST 8 KB box still takes ~1 700 cycles to
  // land, as long as the whole 128 KB box: the landing is latency-, not bandwidth-bound; profiles/r02_gru_pipelined.txt).
  // Groups: B > 64 rows are walked in groups of 64 inside every timestep.  Tuning flag 1 << 27: groups of 32 rows for
  // 32 < B <= 64 (half the landing bytes per group, and with per-group counters the two groups hide each other's
  // handshake; each group still costs a full M = 64 MMA pass).
  const int group_rows = ((a->tuning_flags & (1 << 27)) && B > 32 && B <= GRU_M) ? 32 : GRU_M;
  const int rows = B < group_rows ? B : group_rows;    // rows of one TMA box (B > group_rows: multi-group mode)
  const int groups = (B + group_rows - 1) / group_rows;
  const bool one_box = !(a->tuning_flags & 32) && K % 64 == 0 && B % 8 == 0 && kbc * rows * 128 <= 160 * 1024;
  {
    const uint64_t slots = BWD ? (uint64_t)T : (uint64_t)T + 1;
    const void* base = BWD ? (const void*)a->dgh : (const void*)a->h_ext;
    int rc;
    if (one_box) {
      const uint64_t dims[4] = {64, (uint64_t)B, (uint64_t)(K / 64), slots};
      const uint64_t strides[3] = {(uint64_t)K * 2, 128, (uint64_t)K * 2 * (uint64_t)a->ext_batch};
      const uint32_t box[4] = {64, (uint32_t)rows, (uint32_t)kbc, 1};
      rc = make_tmap_bf16(&tx, base, 4, dims, strides, box, true);
    } else {
      const uint64_t dims[3] = {(uint64_t)K, (uint64_t)B, slots};
      const uint64_t strides[2] = {(uint64_t)K * 2, (uint64_t)K * 2 * (uint64_t)a->ext_batch};
      const uint32_t box[3] = {64, (uint32_t)rows, 1};
      rc = make_tmap_bf16(&tx, base, 3, dims, strides, box, true);
    }
    if (rc) return rc;
  }
  GruParams p{};
  p.batch = B; p.steps = T; p.hidden = H; p.ext_batch = a->ext_batch; p.kbc = kbc;
  p.one_box = one_box ? 1 : 0;
  p.groups = groups;
  p.box_rows = rows;
  p.group_rows = group_rows;
  // Per-group arrival counters: default for FORWARD multi-group launches (measured B=128, T=4000: 6.77 us per timestep
  // against 7.73 with one counter per timestep); the backward kernel keeps one counter per timestep (9.12 against 9.50:
  // its cluster reduction already staggers the groups).  1 << 26 forces one counter per timestep, 1 << 28 per-group
  // counters in the backward kernel too; never combined with the fan-out experiment (which owns the flag lines).
  p.group_sync = (groups > 1 && !(a->tuning_flags & ((1 << 26) | (1 << 24))) && (!BWD || (a->tuning_flags & (1 << 28)))) ? 1 : 0;
  p.hslot = one_box ? rows * 128 : GRU_SLOT;
  p.gi = static_cast<const __nv_bfloat16*>(a->gi);
  p.b_hh = a->b_hh;
  p.h_ext = static_cast<__nv_bfloat16*>(a->h_ext);
  p.hall = static_cast<__nv_bfloat16*>(a->hall);
  p.h_state = a->h_state;
  p.gates = static_cast<__nv_bfloat16*>(a->gates);
  p.dh_out = static_cast<const __nv_bfloat16*>(a->dh_out);
  p.dgi = static_cast<__nv_bfloat16*>(a->dgi);
  p.dgh = static_cast<__nv_bfloat16*>(a->dgh);
  p.dh0 = a->dh0;
  p.c_state = a->c_state;
  p.c_init = a->c_init;
  p.dc0 = a->dc0;
  p.db_ih = a->db_ih;
  p.db_hh = a->db_hh;
  p.sync = a->sync;
  p.flags = a->tuning_flags;
  p.ts = reinterpret_cast<unsigned long long*>(a->debug_ts);

  auto kern = gru_kernel<BWD, C, LSTM, U, MW>;
  SRNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // Sentinel for the fence-free exchange: every slot the kernel will WRITE and then re-read through the grid handshake
  // starts as bf16 NaN (0xFFFF), so an element that has not arrived yet cannot be mistaken for data.  Forward: slots
  // 1..T of h_ext (slot 0 is the caller's h_init); backward: all T slots of dgh.  A single forward timestep never reads
  // what it wrote, and the strict protocol (flag 16) needs no sentinel.
  if (!(a->tuning_flags & 16) && (BWD || T > 1)) {
    const size_t row_bytes = static_cast<size_t>(K) * 2;
    char* base = BWD ? static_cast<char*>(a->dgh) : static_cast<char*>(a->h_ext) + static_cast<size_t>(a->ext_batch) * row_bytes;
    if (a->ext_batch == B)
      SRNN_CUDA(cudaMemsetAsync(base, 0xFF, static_cast<size_t>(T) * B * row_bytes, stream));
    else
      SRNN_CUDA(cudaMemset2DAsync(base, static_cast<size_t>(a->ext_batch) * row_bytes, 0xFF, static_cast<size_t>(B) * row_bytes,
                                  static_cast<size_t>(T), stream));
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(GRU_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[3];
  int na = 0;
  if (T > 1 || C > 1 || (a->tuning_flags & 2)) {        // (flag 2: experiment, force the cooperative launch)
    attrs[na].id = cudaLaunchAttributeCooperative;   // all CTAs co-resident: they spin on one another
    attrs[na].val.cooperative = 1;
    ++na;
  } else if (pdl_enabled()) {
    // one timestep without clusters never waits on another CTA: a plain launch, whose set-up and weight fetch may
    // overlap the previous kernel in the stream
    attrs[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (C > 1) {
    attrs[na].id = cudaLaunchAttributeClusterDimension;
    attrs[na].val.clusterDim.x = C;
    attrs[na].val.clusterDim.y = 1;
    attrs[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attrs;
  cfg.numAttrs = na;
  SRNN_CUDA(cudaLaunchKernelEx(&cfg, kern, tw, tx, p));
  return SRNN_OK;
}

template <bool BWD, int C, bool LSTM, int U>
static int launch_gru(const srnn_gru_args* a, int kbc, cudaStream_t stream) {
  constexpr int MW = gru_mw<BWD, C, LSTM, U>();
  if constexpr (MW != 2 && U == 8 && !LSTM) {          // experiments: debug flag 64 = two issuing warps
    if (a->tuning_flags & 64) return launch_gru_mw<BWD, C, LSTM, U, 2>(a, kbc, stream);
  }
  return launch_gru_mw<BWD, C, LSTM, U, MW>(a, kbc, stream);
}

// The occupancy answer behind pick_cluster() depends on the device and on H only: remembered per (device, hidden),
// guarded by a mutex so that several host threads (one per device, or several on one) may call concurrently.
struct PickCache {
  std::mutex mu;
  int hidden[kMaxDevices];
  int c[kMaxDevices];
  int kbc[kMaxDevices];
  PickCache() { for (int i = 0; i < kMaxDevices; ++i) hidden[i] = -1; }
  int get(int h, int* kbc_out) {                       // -> cluster size (0 = nothing fits), or -1 if unknown
    const int d = current_device();
    std::lock_guard<std::mutex> lock(mu);
    if (hidden[d] != h) return -1;
    *kbc_out = kbc[d];
    return c[d];
  }
  void put(int h, int c_, int kbc_) {
    const int d = current_device();
    std::lock_guard<std::mutex> lock(mu);
    hidden[d] = h; c[d] = c_; kbc[d] = kbc_;
  }
};

// Cluster size: K split in whole K blocks, units tile H, slice fits smem, all clusters co-resident.
template <bool BWD, bool LSTM, int U>
static int pick_cluster(int H, int* kbc_out) {
  const int K = BWD ? (LSTM ? 4 : 3) * H : H;
  const int kb_total = (K + 63) / 64;
  // measured (B=64, H=1024, U=8), us per step forward: C=1 (whole K in one CTA, no cluster exchange, 128 KB
  // of h per CTA and step) 3.9-4.1, C=2 4.4-4.5, C=4 5.6 (exchange volume grows with C);
  // backward needs C>=4 for the 3H-wide slice to fit shared memory
  const int fwd_order[4] = {1, 2, 4, 8};
  const int bwd_order[4] = {4, 8, 2, 1};
  for (int i = 0; i < 4; ++i) {
    const int c = BWD ? bwd_order[i] : fwd_order[i];
    if (kb_total % c != 0 || H % (U * c) != 0) continue;
    const int kbc = kb_total / c;
    if (kbc > GRU_MAX_KBC) continue;
    bool fits = false;
    if constexpr (U == 8) {
      switch (c) {
        case 8: fits = gru_fits<BWD, 8, LSTM, 8>(H, kbc); break;
        case 4: fits = gru_fits<BWD, 4, LSTM, 8>(H, kbc); break;
        case 2: fits = gru_fits<BWD, 2, LSTM, 8>(H, kbc); break;
        default: fits = gru_fits<BWD, 1, LSTM, 8>(H, kbc); break;
      }
    } else {                                           // wide CTAs: clusters of 2 or 4 only (MMA N <= 256)
      switch (c) {
        case 4: fits = gru_fits<BWD, 4, LSTM, 16>(H, kbc); break;
        case 2: fits = gru_fits<BWD, 2, LSTM, 16>(H, kbc); break;
        default: fits = false; break;
      }
    }
    if (!fits) continue;
    *kbc_out = kbc;
    return c;
  }
  return 0;
}

template <bool BWD, bool LSTM, int U>
static int dispatch_gru_u(const srnn_gru_args* a, cudaStream_t stream) {
  static PickCache cache;                              // per instantiation; keyed by (device, hidden)
  int kbc = 0;
  int c = cache.get(a->hidden, &kbc);
  if (c < 0) {
    c = pick_cluster<BWD, LSTM, U>(a->hidden, &kbc);
    cache.put(a->hidden, c, kbc);
  }
  if ((a->tuning_flags >> 8) & 15) {                  // experiments: force a cluster size (bits 8-11)
    const int forced = (a->tuning_flags >> 8) & 15;
    const int kb_total = ((BWD ? (LSTM ? 4 : 3) : 1) * a->hidden + 63) / 64;
    if (kb_total % forced == 0 && a->hidden % (U * forced) == 0 && kb_total / forced <= GRU_MAX_KBC) {
      c = forced;
      kbc = kb_total / forced;
    }
  }
  SRNN_CHECK_ARG(c > 0, "gru: no cluster decomposition fits hidden=%d with %d units per CTA", a->hidden, U);
  if constexpr (U == 8) {
    switch (c) {
      case 8: return launch_gru<BWD, 8, LSTM, 8>(a, kbc, stream);
      case 4: return launch_gru<BWD, 4, LSTM, 8>(a, kbc, stream);
      case 2: return launch_gru<BWD, 2, LSTM, 8>(a, kbc, stream);
      default: return launch_gru<BWD, 1, LSTM, 8>(a, kbc, stream);
    }
  } else {
    switch (c) {
      case 4: return launch_gru<BWD, 4, LSTM, 16>(a, kbc, stream);
      default: return launch_gru<BWD, 2, LSTM, 16>(a, kbc, stream);
    }
  }
}

template <bool BWD, bool LSTM>
static int dispatch_gru(const srnn_gru_args* a, cudaStream_t stream) {
  // 16 units per CTA when asked for (and possible), else 8; more than the SM count of CTAs never fits
  bool wide = a->units_per_cta == 16 && a->hidden % 32 == 0;
  if (wide) {                                          // fall back to 8 units when the wide slice does not fit
    static PickCache probe;
    int kbc = 0;
    int probed_c = probe.get(a->hidden, &kbc);
    if (probed_c < 0) {
      probed_c = pick_cluster<BWD, LSTM, 16>(a->hidden, &kbc);
      probe.put(a->hidden, probed_c, kbc);
    }
    wide = probed_c > 0;
  }
  if (!wide) {
    SRNN_CHECK_ARG(a->hidden / 8 <= sm_count(), "gru: hidden/8 = %d CTAs exceeds the SM count %d", a->hidden / 8,
                   sm_count());
    return dispatch_gru_u<BWD, LSTM, 8>(a, stream);
  }
  return dispatch_gru_u<BWD, LSTM, 16>(a, stream);
}

}  // namespace srnn

using namespace srnn;

static int check_common(const srnn_gru_args* a) {
  SRNN_CHECK_ARG(a != nullptr, "gru: null args");
  SRNN_CHECK_ARG(a->batch > 0 && a->batch <= 8 * GRU_M, "gru: batch must be in 1..512 (got %d)", a->batch);
  SRNN_CHECK_ARG(a->steps > 0 && a->hidden > 0 && a->hidden % 8 == 0, "gru: bad steps/hidden (%d, %d)", a->steps,
                 a->hidden);
  SRNN_CHECK_ARG(a->w_hh && a->h_ext && a->gates && a->sync, "gru: null buffer");
  SRNN_CHECK_ARG(a->ext_batch >= a->batch, "gru: ext_batch (%d) must be >= batch (%d)", a->ext_batch, a->batch);
  SRNN_CHECK_ARG(a->cell == 0 || a->cell == 1, "gru: cell must be 0 (GRU) or 1 (LSTM)");
  SRNN_CHECK_ARG(a->units_per_cta == 0 || a->units_per_cta == 8 || a->units_per_cta == 16,
                 "gru: units_per_cta must be 0, 8 or 16");
#ifndef SRNN_DEBUG
  SRNN_CHECK_ARG((a->tuning_flags & (1 | 4)) == 0,
                 "gru: tuning_flags 1 and 4 select instrumented modes that exist only in -DSRNN_DEBUG builds");
#endif
  return SRNN_OK;
}

extern "C" int srnn_gru_forward(const srnn_gru_args* a, srnn_stream_t stream) {
  int rc = check_common(a);
  if (rc) return rc;

  SRNN_CHECK_ARG(a->gi && a->b_hh && a->h_state, "gru_forward: null buffer");
  if (a->cell == 1) {
    SRNN_CHECK_ARG(a->c_state, "lstm forward: c_state required");
    return dispatch_gru<false, true>(a, static_cast<cudaStream_t>(stream));
  }
  return dispatch_gru<false, false>(a, static_cast<cudaStream_t>(stream));
}

extern "C" int srnn_gru_backward(const srnn_gru_args* a, srnn_stream_t stream) {
  int rc = check_common(a);
  if (rc) return rc;

  SRNN_CHECK_ARG(a->dh_out && a->dgi && a->dgh && a->dh0, "gru_backward: null buffer");
  if (a->cell == 1) {
    SRNN_CHECK_ARG(a->c_init && a->dc0, "lstm backward: c_init and dc0 required");
    return dispatch_gru<true, true>(a, static_cast<cudaStream_t>(stream));
  }
  return dispatch_gru<true, false>(a, static_cast<cudaStream_t>(stream));
}
