#!/usr/bin/env python
"""Benchmark of the teacher-forced SampleRNN training step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A *step* is one training step on one sequential-loader chunk: forward + NLL + backward
(+ gradient all-reduce for N > 1) + AdamClipped.  The workload is BASELINE config 2: 3-tier
SampleRNN GRU (ratios [4,4] = frame sizes 16/4, H=1024), 64 utterance slots per GPU x chunks of
L=1000 frames (RF = 16 000 samples = 1 s of 16 kHz audio), hidden state carried across chunks; an
8 s utterance batch is 8 consecutive steps.  Weak scaling: every rank owns 64 slots.

Printed JSON (one line, rank 0):
  value     audio samples/s over all GPUs, inputs already resident in HBM, no host sync per step
  e2e       the same metric through the public module API with pinned HOST buffers: H2D copies of
            x / y / conds and the D2H read of the loss are inside the timed region
  roofline  the DOMINANT kernel of the step - the persistent recurrent kernel - timed live with CUDA events: us per
            timestep, share of the step and achieved algorithmic HBM rate vs the measured copy peak
  roofline_gemm  the largest contraction (comb_layer forward, tcgen05) vs the measured sustained and burst bf16 peaks
  recurrence     all recurrent launches of the step (direction x timesteps)
  cpu_baseline   the CPU oracle port timed on this box's host cores on a bounded sample of the same workload
  gpu_eager_baseline  informational: the oracle port (plain torch ops) on the same GPU, fp32 and bf16 autocast

``--impl reference`` times the reference's CPU path (the oracle port; the reference is pure Python
and cannot travel to the GPU box) on the host cores with the same metric/unit.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RATIOS, LAYERS, HIDDEN = [4, 4], [1, 1], [1024, 1024]
SLOTS_PER_GPU, SEQ_LEN = 64, 1000
N_SPEAKERS = 126
METRIC = 'teacher-forced training audio samples/sec'
WORKLOAD = ('config2: 3-tier SampleRNN GRU ratios [4,4] H=1024, 64 slots/GPU x 1 s chunks (L=1000, RF=16000) '
            'of 8 s utterances with hidden-state carry, acoustic conds U=43, 126 speakers')


CHUNKS = 8                   # sequential-loader chunks per utterance batch (8 s of audio)
EXTRA = {}                   # extension keywords of SampleRNNModel (rnn_cell) for the other BASELINE configs
SPEAKER = ('embedding', 15)  # conds_speaker_type, conds_speaker_size


def select_workload(name):
    """BASELINE.json configs[1] is the default and the one the driver measures; configs[2] and configs[3] can be
    timed with ``--workload`` at their PER-GPU shard (weak scaling: the full configs are 8 such ranks)."""
    global RATIOS, LAYERS, HIDDEN, SLOTS_PER_GPU, CHUNKS, EXTRA, SPEAKER, WORKLOAD
    if name == 'config1':
        RATIOS, SLOTS_PER_GPU, CHUNKS = [20, 4], 8, 1
        WORKLOAD = ('config1: 2-tier SampleRNN of config.default.json (ratios [20,4], H=1024, 47.8 M parameters), '
                    '8 slots x one 1 s chunk (L=200, RF=16000), acoustic conds U=43, 126 speakers')
    elif name == 'config3':
        SLOTS_PER_GPU, EXTRA, SPEAKER = 16, dict(rnn_cell='lstm'), ('pase', 100)
        WORKLOAD = ('config3: 3-tier SampleRNN LSTM ratios [4,4] H=1024 with a 100-d PASE speaker vector, 16 slots/GPU '
                    '(global batch 128 on 8 GPUs) x 1 s chunks of 8 s utterances with (h, c) carry')
    elif name == 'config4':
        RATIOS, LAYERS, HIDDEN, SLOTS_PER_GPU, CHUNKS = [4, 4, 4], [1, 1, 1], [1024, 1024, 1024], 32, 16
        WORKLOAD = ('config4: 4-tier SampleRNN GRU ratios [4,4,4] H=1024, 32 slots/GPU (global batch 256 on 8 GPUs) x '
                    '16 sequential 1 s chunks (L=250, RF=16000) of 16 s utterances with hidden-state carry')
    elif name != 'config2':
        raise SystemExit(f'unknown workload {name}')


def model_kwargs(seq_len=None):
    if seq_len is None:
        seq_len = seq_len_default()
    return dict(conds_speaker_type=SPEAKER[0], conds_speaker_n=N_SPEAKERS, conds_speaker_size=SPEAKER[1],
                conds_utterance_type='acoustic', conds_utterance_linguistic_n=[9, 5, 4, 3],
                conds_utterance_linguistic_emb_size=10, conds_size=50, sequence_length=seq_len, ratios=RATIOS,
                rnn_layers=LAYERS, rnn_hidden_size=HIDDEN, q_type_ulaw=True, q_levels=256, **EXTRA)


def seq_len_default():
    """Frames of the top tier per chunk such that a chunk is 16 000 samples (1 s of 16 kHz audio)."""
    fs = 1
    for r in RATIOS:
        fs *= r
    return 16000 // fs


def flops_per_sample_fwd():
    """dense model FLOPs of the reference arithmetic per audio sample."""
    h, c, q = HIDDEN[0], 50, 256
    r0 = RATIOS[0]
    f = 2 * (r0 * q * h + c * h + 3 * h * h + h * h + h * q)
    fs = 1
    for n, r in enumerate(RATIOS):
        fs *= r
        gates = 4 if EXTRA.get('rnn_cell') == 'lstm' else 3
        f += (1.0 / fs) * 2 * (fs * h + c * h + LAYERS[n] * 2 * gates * h * h + r * h * h)
    return f


# --------------------------------------------------------------------------------------------------
# CPU reference arm (oracle port) and the informational GPU-eager leg
# --------------------------------------------------------------------------------------------------
CPU_SAMPLE_SLOTS = 8         # slots of the workload's chunk the CPU arm runs per step (a bounded sample: same model, same
                             # chunk length L, carry across chunks; per-sample CPU cost does not depend on the slot count)


def config_dict(world=1, slots=None):
    """The workload description BOTH arms print (same keys, same values)."""
    slots = SLOTS_PER_GPU if slots is None else slots
    fs = 1
    for r in RATIOS:
        fs *= r
    return dict(workload=WORKLOAD, ratios=RATIOS, hidden=HIDDEN[0], seq_len=seq_len_default(),
                samples_per_chunk=seq_len_default() * fs, chunks_per_utterance=CHUNKS, slots_per_gpu=slots,
                cell=EXTRA.get('rnn_cell', 'gru'))


def reference_trainer(device, batch, steps, warmup, autocast=False):
    """The oracle port (plain torch ops, the reference's operator choice: fused GRU, Conv1d, ConvTranspose1d, Linear;
    fp32, default TF32 flags) running the workload's model on ``batch`` slots of its chunk (same L) on ``device``.
    step = forward + NLL + backward + AdamClipped, hidden state carried.  -> dict(value, ms_per_step, steps)."""
    from oracle import samplernn_oracle as O
    seq_len = seq_len_default()
    spec = O.ModelSpec(RATIOS, LAYERS, HIDDEN, seq_len, cell=EXTRA.get('rnn_cell', 'gru'))
    params = {k: v.to(device) for k, v in O.init_params(spec, conds_speaker_n=N_SPEAKERS).items()}
    trainer = O.CpuTrainer(spec, params)
    wav, conds, spk = O.synthetic_utterances(spec, batch, warmup + steps)
    wav, conds, spk = wav.to(device), conds.to(device), spk.to(device)
    cuda = torch.device(device).type == 'cuda'
    times = []
    for k in range(warmup + steps):
        x, y, c = O.chunk_of(spec, wav, conds, k)
        if cuda:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        if autocast:
            with torch.autocast('cuda', dtype=torch.bfloat16):
                trainer.step(x, y, c, spk, [1] * batch if k == 0 else [0] * batch)
        else:
            trainer.step(x, y, c, spk, [1] * batch if k == 0 else [0] * batch)
        if cuda:
            torch.cuda.synchronize()
        if k >= warmup:
            times.append(time.perf_counter() - t0)
    per_step = batch * spec.receptive_field
    total = sum(times)
    return dict(value=per_step * len(times) / total, ms_per_step=1e3 * total / len(times), steps=len(times),
                rf=spec.receptive_field, seq_len=seq_len)


def cpu_reference(steps, warmup, batch=CPU_SAMPLE_SLOTS):
    torch.set_num_threads(os.cpu_count() or 1)
    batch = min(batch, SLOTS_PER_GPU)
    r = reference_trainer('cpu', batch, steps, warmup)
    r['cores'] = torch.get_num_threads()
    r['sample'] = (f'oracle port (torch CPU fp32, the reference\'s operators), the workload\'s model on {batch} of its '
                   f'{SLOTS_PER_GPU} slots x its own chunk length L={r["seq_len"]} (RF={r["rf"]}) with carry, '
                   f'{warmup} warm-up + {r["steps"]} timed steps of {batch * r["rf"]} samples')
    return r


def gpu_eager_baseline(dev, slots):
    """Informational: the same oracle port moved to the GPU (``.to('cuda')`` is all the reference does, runner.py:27,50):
    fp32, torch's default flags (matmul TF32 off, cuDNN TF32 on), cuDNN GRU, the full per-GPU chunk of the workload.
    This is the 'reference-GPU-eager' denominator of BASELINE.json's target; it runs after every timed region."""
    out = {}
    for key, autocast in (('fp32', False), ('bf16_autocast', True)):
        try:
            torch.cuda.empty_cache()
            r = reference_trainer(dev, slots, 2, 1, autocast=autocast)
            out[key] = dict(value=r['value'], unit='samples/s', ms_per_step=r['ms_per_step'], steps=r['steps'],
                            slots=slots)
        except Exception as e:                                          # noqa: BLE001  (informational leg only)
            out[key] = dict(unavailable=f'{type(e).__name__}: {str(e)[:160]}')
    out['what'] = ('oracle port of the reference (plain torch ops: cuDNN GRU / Conv1d / ConvTranspose1d, cuBLAS Linear) on '
                   'cuda:0, default torch flags, same chunk and slot count as the GPU arm, 1 warm-up + 2 timed steps; '
                   'kind = port (the reference itself cannot travel to the GPU box)')
    return out


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    r = cpu_reference(max(1, min(args.steps, 3)), max(1, min(args.warmup, 1)))
    line = dict(impl='reference', metric=METRIC, value=r['value'], unit='samples/s', n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=r['ms_per_step'], higher_is_better=True, scaling=scaling_kind(args),
                vs_baseline=None, dtype='f32', data='synthetic', config=config_dict(args.gpus, slots_for(args)),
                cpu_baseline=dict(value=r['value'], unit='samples/s', cores=r['cores'], kind='port', sample=r['sample']),
                e2e=dict(value=r['value'], unit='samples/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def scaling_kind(args):
    return 'strong' if args.global_batch else 'weak'


def slots_for(args):
    if args.global_batch:
        if args.global_batch % args.gpus:
            raise SystemExit(f'--global-batch {args.global_batch} is not divisible by --gpus {args.gpus}')
        return args.global_batch // args.gpus
    return SLOTS_PER_GPU


# --------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
              'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), f'--query-gpu={self.FIELDS}',
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        self.proc.terminate()
        out = self.proc.communicate()[0]
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in out.strip().splitlines():
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[2:6]):
                if flag.lower().startswith('active'):
                    reasons.add(name)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def recurrence_report(events, steps_timed, step_ms, b, hidden, lstm, peak_gbs):
    """Per recurrent launch kind (direction x timesteps): time inside the step, us per timestep, share of the step and
    the achieved algorithmic HBM rate.  Algorithmic bytes per (slot, timestep), bf16:
      forward : read the 3H (4H LSTM) input pre-activations + write H outputs                      = 8 KB at H=1024
      backward: read dL/dh (H) + the saved gates (4H; 5H LSTM) + h_{t-1} (H; LSTM c_{t-1} instead),
                write both gate-gradient copies (2 x 3H; 2 x 4H LSTM)                                 = 24 KB at H=1024."""
    g = 4 if lstm else 3
    per = {'fwd': 2.0 * (g * hidden + hidden), 'bwd': 2.0 * (hidden + (g + 1) * hidden + hidden + 2 * g * hidden)}
    out = []
    for tag, evs in sorted(events.items()):
        if not tag.startswith('rnn_'):
            continue
        kind, t = tag[4:7], int(tag.split('_T')[1])
        f32 = '_f32_' in tag                                   # --precision fp32: one split-operand GEMM + one cell kernel per timestep
        ms = [a.elapsed_time(b_) for a, b_ in evs]
        per_step_ms = sum(ms) / steps_timed                    # all launches of this kind in one training step
        groups = max(1, round(len(ms) / steps_timed))          # slot groups (batch > 64) run one after the other
        rows = min(b, 64 * groups) if groups > 1 else b
        bytes_launch = per[kind] * (b / groups) * t * (2 if f32 else 1)
        avg = sum(ms) / len(ms)
        name = f'gru_f32 host loop <{kind}> T={t}' if f32 else f'gru_kernel<{kind}> T={t}' + (' (LSTM)' if lstm else '')
        out.append(dict(kernel=name, launches_per_step=groups,
                        avg_launch_ms=avg, us_per_timestep=1e3 * avg / t, ms_per_step=per_step_ms,
                        share_of_step=per_step_ms / step_ms, algorithmic_bytes_per_launch=bytes_launch,
                        achieved_gbs=bytes_launch / (avg * 1e-3) / 1e9, frac_of_hbm_peak=bytes_launch / (avg * 1e-3) / 1e9 / peak_gbs))
    return sorted(out, key=lambda d: -d['ms_per_step'])


def run_gpu(args):
    import torch.distributed as dist
    from samplernn_pase_b200 import SampleRNNModel, ops, synthetic
    from samplernn_pase_b200.parallel import DataParallelTrainer

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus:
        raise SystemExit(f'--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run for N > 1')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    torch.manual_seed(1234)
    model = SampleRNNModel(fused_loss=True, precision=args.precision, **model_kwargs()).to(dev)
    trainer = DataParallelTrainer(model, lr=1e-4)
    fs = int(model.frame_size)
    rf = int(model.receptive_field)
    b = slots_for(args)
    chunks = CHUNKS
    seq_len = seq_len_default()
    wav, conds, spk = synthetic.synthetic_utterances(fs, rf, seq_len, b, chunks, seed=4321 + rank, n_speakers=N_SPEAKERS)
    info = [{'speaker': {'index': int(s)}} for s in spk]
    if SPEAKER[0] == 'pase':
        vecs = torch.randn(b, SPEAKER[1], generator=torch.Generator().manual_seed(99 + rank))
        info = [{'speaker': {'pase': vecs[i]}} for i in range(b)]
    host = []
    for k in range(chunks):
        x, y, c = synthetic.chunk_of(fs, rf, seq_len, wav, conds, k)
        host.append((x.pin_memory(), y.pin_memory(), c.pin_memory()))
    resident = [tuple(t.to(dev) for t in h) for h in host]
    resets = [torch.ones(b, dtype=torch.int64), torch.zeros(b, dtype=torch.int64)]
    global_rows = b * rf * world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def loop(n_steps, e2e, first):
        last = None
        for s in range(first, first + n_steps):
            k = s % chunks
            reset = resets[0] if k == 0 else resets[1]
            if e2e:
                x, y, c = (t.to(dev, non_blocking=True) for t in host[k])
                loss, _ = trainer.step(x, y, c, info, reset)            # reads the loss back (D2H)
                last = loss
            else:
                x, y, c = resident[k]
                last, _ = trainer.step(x, y, c, info, reset, global_count=global_rows)
        return last

    def timed(n_steps, e2e, first):
        barrier()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        last = loop(n_steps, e2e, first)
        t1.record()
        barrier()
        own = t0.elapsed_time(t1)
        ms = torch.tensor([own], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms), float(last), own

    loop(args.warmup, False, 0)
    sampler = ClockSampler(local) if rank == 0 else None
    ops.launch_count = 0
    ops.event_log = {}
    ms, loss, own_ms = timed(args.steps, False, args.warmup)
    launches = ops.launch_count
    events = ops.event_log
    ops.event_log = None
    clocks = sampler.stop() if sampler else None
    kernel_ms = [a.elapsed_time(b_) for a, b_ in events.get('comb_layer_fwd', [])]
    loop(1, True, args.warmup + args.steps)
    ms_e2e, loss_e2e, _ = timed(args.steps, True, args.warmup + args.steps + 1)

    # data-parallel consistency: every rank must hold bit-identical parameters after the timed steps, and the per-rank
    # step times show whether the max-over-ranks is one slow rank (clock skew) or all of them
    ranks_equal, rank_ms = None, None
    if world > 1:
        chk = trainer.flat.flat_param.view(torch.int32).to(torch.int64).sum().reshape(1)
        gathered = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(gathered, chk)
        ranks_equal = all(int(g) == int(gathered[0]) for g in gathered)
        t_own = torch.tensor([own_ms / args.steps], device=dev)
        t_all = [torch.zeros_like(t_own) for _ in range(world)]
        dist.all_gather(t_all, t_own)
        rank_ms = [round(float(t), 3) for t in t_all]
        if not ranks_equal:
            raise SystemExit(f'rank parameter checksums differ after the timed steps: {[int(g) for g in gathered]}')

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except OSError:
            pass
        # fallbacks stated in /opt/skills/guides/B200_PROFILING.md when the driver-written file is absent
        peak_tf = peaks.get('bf16_tflops_sustained', 1400.0)
        peak_tf_burst = peaks.get('bf16_tflops', 1700.0)
        peak_gbs = peaks.get('hbm_gbs', 6500.0)
        m_rows, h = b * rf, HIDDEN[0]
        r0 = RATIOS[0]
        step_ms = ms / args.steps
        total = args.steps * global_rows
        f_fwd = flops_per_sample_fwd()
        lstm = EXTRA.get('rnn_cell') == 'lstm'
        rec = recurrence_report(events, args.steps, own_ms / args.steps, b, h, lstm, peak_gbs)
        # traffic: dram__bytes_read.sum + dram__bytes_write.sum of one launch from an `ncu --set full` capture AT THE BENCHED
        # SIZE (profiles/r02_ncu_traffic.json, written by scripts/ncu_traffic.py from the capture); null if not captured
        ncu = {}
        try:
            ncu = json.load(open(os.path.join(ROOT, 'profiles', 'r02_ncu_traffic.json')))
        except (OSError, ValueError):
            pass
        if args.workload == 'config2' and not args.global_batch:
            for r in rec:                                 # measured DRAM bytes per launch where a capture exists
                r['traffic'] = ncu.get(r['kernel'])
        top = rec[0] if rec else None
        roofline = None
        if top:
            roofline = dict(bound='hbm', kernel=top['kernel'] + f', {b} slots, H={h}: the dominant kernel '
                            f'({100 * top["share_of_step"]:.1f} % of the step; all recurrent launches together '
                            f'{100 * sum(r["share_of_step"] for r in rec):.1f} %)',
                            achieved=top['achieved_gbs'], peak=peak_gbs, unit='GB/s', frac=top['frac_of_hbm_peak'],
                            traffic=ncu.get(top['kernel'].split(' (')[0]) if args.workload == 'config2' and not args.global_batch else None,
                            traffic_algorithmic=top['algorithmic_bytes_per_launch'], avg_launch_ms=top['avg_launch_ms'],
                            us_per_timestep=top['us_per_timestep'], share_of_step=top['share_of_step'],
                            note='latency-bound by construction (one grid-wide exchange of h per timestep): us_per_timestep is '
                                 'the figure of merit, the HBM fraction shows how far from bandwidth-bound it is',
                            peak_source='MEASURED_PEAKS.json hbm_gbs' if peaks else 'B200_PROFILING.md fallback')
        kc = r0 * 256 + h                                         # comb_layer forward GEMM as executed: A = [one-hot windows | upper]
        gemm_flops = 2.0 * m_rows * h * kc
        avg_ms = sum(kernel_ms) / max(len(kernel_ms), 1)
        achieved = gemm_flops / (avg_ms * 1e-3) / 1e12 if avg_ms else None
        roofline_gemm = dict(
            bound='tensor', kernel=f'gemm_kernel<256,NT,frame-term+relu epilogue> (comb_layer forward, m x {h} x {kc}: '
                                   f'A = [one-hot windows | upper], tcgen05)',
            achieved=achieved, peak=peak_tf, unit='TFLOP/s', frac=(achieved / peak_tf) if achieved else None,
            frac_of_burst_peak=(achieved / peak_tf_burst) if achieved else None, peak_burst=peak_tf_burst,
            traffic=ncu.get('comb_layer_fwd') if args.workload == 'config2' and not args.global_batch else None,
            traffic_algorithmic=512.0 * m_rows + 2.0 * m_rows * 2 * h + 2.0 * (m_rows // r0) * h + 2.0 * h * kc,
            launches_timed=len(kernel_ms), avg_launch_ms=avg_ms, share_of_step=avg_ms / step_ms if avg_ms else None,
            note='half of the A operand is one-hot (1 non-zero in 256): it draws less power than the dense random operands '
                 'the peaks were measured with, so the sustained-relative fraction can exceed 1; the burst-relative one cannot')
        del trainer, model, resident
        # host-side and eager baselines: rank 0 at N = 1 only (they describe one GPU / one socket)
        cpu = cpu_reference(2, 1) if world == 1 and not args.no_cpu else None
        eager = gpu_eager_baseline(dev, b) if world == 1 and not args.no_eager else None
        line = dict(
            metric=METRIC, value=total / (ms * 1e-3), unit='samples/s', n_gpus=world, steps=args.steps,
            warmup=args.warmup, ms_per_step=step_ms, higher_is_better=True, scaling=scaling_kind(args), vs_baseline=None,
            dtype='bf16' if args.precision == 'bf16' else 'f32',
            data='synthetic',
            config=config_dict(world, b),
            run=dict(samples_per_step_per_gpu=b * rf, parallelism=f'dp{world}', global_batch=b * world,
                     l2_policy='inputs and activations per step (>10 GB) exceed the 126 MB L2; no flush needed',
                     loss_mode='fused log-softmax+NLL epilogue' if args.precision == 'bf16' else 'fp32 log-softmax rows',
                     arithmetic='bf16 operands, fp32 accumulation' if args.precision == 'bf16' else
                     'fp32-tolerance mode: fp32 activations, split-bf16 (hi+lo) operands into the tcgen05 GEMM (3x K), '
                     'recurrence as one GEMM + one cell kernel per timestep; NOT the headline configuration',
                     final_loss=loss,
                     model_train_mflop_per_sample=3 * f_fwd / 1e6,
                     model_tflops_achieved=3 * f_fwd * total / (ms * 1e-3) / 1e12,
                     model_tflops_frac_of_sustained_peak=3 * f_fwd * total / (ms * 1e-3) / 1e12 / peak_tf / world,
                     rank_params_identical=ranks_equal, rank_ms_per_step=rank_ms),
            clocks=clocks,
            e2e=dict(value=total / (ms_e2e * 1e-3), unit='samples/s', ms_per_step=ms_e2e / args.steps,
                     h2d_bytes_per_step=sum(t.numel() * t.element_size() for t in host[0]), d2h_bytes_per_step=8,
                     final_loss=loss_e2e),
            gpu_launches=launches,
            roofline=roofline, roofline_gemm=roofline_gemm, recurrence=rec,
            cpu_baseline=dict(value=cpu['value'], unit='samples/s', cores=cpu['cores'], kind='port',
                              sample=cpu['sample']) if cpu else None,
            gpu_eager_baseline=eager,
        )
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_generation(args):
    """BASELINE configs[4]: batched autoregressive generation, 256 utterances x 4 s (64 000 samples) with the config-2 model.
    step = one ``SampleRNNModel.test`` call (model.py:289-351) for all utterances; value = generated samples / s."""
    from samplernn_pase_b200 import SampleRNNModel, ops
    if args.gpus != 1:
        raise SystemExit('config5 is a single-GPU workload (replicas only: utterances are independent)')
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(0)
    torch.manual_seed(1234)
    utts, frames = 256, args.gen_frames
    model = SampleRNNModel(**model_kwargs(frames)).to(dev)
    fs = int(model.frame_size)
    host = torch.randn(utts, frames, 43, generator=torch.Generator().manual_seed(4321)).pin_memory()
    info = [{'speaker': {'index': i % N_SPEAKERS}} for i in range(utts)]
    resident = host.to(dev)
    gen = torch.Generator(device=dev).manual_seed(7)
    steps, warmup = args.steps, max(args.warmup, 3)

    def timed(n, e2e):
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(n):
            y = model.test(host.to(dev, non_blocking=True) if e2e else resident, info, generator=gen)
            if e2e:
                y = y.cpu()
        t1.record()
        torch.cuda.synchronize()
        return t0.elapsed_time(t1), y

    timed(warmup, False)
    sampler = ClockSampler(0)
    ops.launch_count = 0
    ms, y = timed(steps, False)
    launches = ops.launch_count
    clocks = sampler.stop()
    ms_e2e, y = timed(steps, True)
    assert y.shape == (utts, (frames + 1) * fs) and int(y.min()) >= 0 and int(y.max()) <= 255
    assert len(set(y[0, fs:].tolist())) > 8                                     # it really samples
    total = steps * utts * frames * fs
    line = dict(metric='batched autoregressive generation audio samples/sec', value=total / (ms * 1e-3), unit='samples/s',
                n_gpus=1, steps=steps, warmup=warmup, ms_per_step=ms / steps, higher_is_better=True, scaling='weak',
                vs_baseline=None, dtype='bf16', data='synthetic',
                config=dict(workload=f'config5: batched autoregressive generation, {utts} utterances x {frames * fs} samples '
                                     f'({frames * fs / 16000:.2f} s of 16 kHz audio), config-2 model (ratios [4,4], H=1024), '
                                     'device-side Philox multinomial draws', utterances=utts, samples_per_utterance=frames * fs),
                run=dict(us_per_sample_step=1e3 * ms / steps / (frames * fs),
                         note='one sample step = all utterances advance by one sample; the FS step programs of a frame are '
                              'one CUDA graph (captured inside every call: set-up + capture are part of the timed step)'),
                clocks=clocks,
                e2e=dict(value=total / (ms_e2e * 1e-3), unit='samples/s', ms_per_step=ms_e2e / steps,
                         h2d_bytes_per_step=host.numel() * 4, d2h_bytes_per_step=utts * (frames + 1) * fs * 8),
                gpu_launches=launches, roofline=None, cpu_baseline=None)
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=8)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='config2', choices=['config1', 'config2', 'config3', 'config4', 'config5'])
    ap.add_argument('--gen-frames', type=int, default=4000, help='config5: top-tier frames per utterance (4000 = 4 s)')
    ap.add_argument('--global-batch', type=int, default=0,
                    help='fix the GLOBAL number of slots (strong scaling: slots per GPU = global / N) instead of the '
                         'per-GPU slot count of the workload')
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'],
                    help="fp32: the fp32-tolerance arithmetic mode (validation; ~4x slower; GRU workloads only)")
    ap.add_argument('--no-eager', action='store_true', help='skip the informational GPU-eager leg')
    ap.add_argument('--no-cpu', action='store_true', help='skip the CPU baseline leg (profiling runs)')
    args = ap.parse_args()
    if args.workload == 'config5':
        if args.impl == 'reference':
            print(json.dumps(dict(impl='reference', unavailable='the reference generates one utterance per call in a Python '
                                  'per-sample loop (model.py:289-351); no batched CPU generation arm is timed')))
            return
        args.steps = min(args.steps, 3)
        return run_generation(args)
    select_workload(args.workload)
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
