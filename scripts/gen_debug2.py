import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samplernn_pase_b200 import SampleRNNModel
from oracle import samplernn_oracle as O
spec = O.ModelSpec([2, 2, 3], [1, 2, 1], [64, 64, 64], 3)
params = O.init_params(spec, conds_speaker_n=5, perturb=0.1)
model = SampleRNNModel('embedding', 5, 15, 'acoustic', [9, 5, 4, 3], 10, 50, 3, [2, 2, 3], [1, 2, 1], [64, 64, 64], True, 256).cuda()
model.load_state_dict(params)
bsz, fs = 3, 12
info = [{'speaker': {'index': i}} for i in range(bsz)]
utt1 = torch.randn(1, 3, 43)
print('single', model.test(utt1.cuda(), info[0]).shape)
torch.cuda.synchronize()
for k, v in model.state_dict().items():
    dmax = float((v.cpu() - params[k]).abs().max())
    if dmax > 0: print('PARAM CHANGED', k, dmax)
print('single eager', model.test(utt1.cuda(), info[0], use_graphs=False).shape)
for k, v in model.state_dict().items():
    dmax = float((v.cpu() - params[k]).abs().max())
    if dmax > 0: print('PARAM CHANGED after eager', k, dmax)
for t2, graphs in ((5, True), (5, False), (5, True), (8, True)):
    utt2 = torch.randn(bsz, t2, 43, generator=torch.Generator().manual_seed(4))
    torch.cuda.manual_seed(11)
    y2, logp2 = model.test(utt2.cuda(), info, return_logp=True, use_graphs=graphs)
    y2 = y2.cpu()
    spec2 = O.ModelSpec([2, 2, 3], [1, 2, 1], [64, 64, 64], t2)
    rf2 = t2 * fs
    ref2 = O.forward_indices(params, spec2, y2[:, :rf2 + fs - 1], y2[:, fs:fs + rf2], utt2, torch.arange(bsz), [1] * bsz)[0]
    d = (logp2.cpu() - ref2).abs().amax(dim=(0, 2))
    print('t', t2, 'graphs', graphs, 'max', float(d.max()), 'per frame', [round(float(d[f*fs:(f+1)*fs].max()), 3) for f in range(t2)])
