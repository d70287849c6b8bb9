import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samplernn_pase_b200 import SampleRNNModel, generate
torch.manual_seed(0)
model = SampleRNNModel('embedding', 5, 15, 'acoustic', [9, 5, 4, 3], 10, 50, 3, [2, 2, 3], [1, 2, 1], [64, 64, 64], True, 256).cuda()
with torch.no_grad():
    for p in model.parameters():
        if float(p.abs().max()) == 0: p.add_(0.1 * torch.randn_like(p))
utt = torch.randn(3, 5, 43).cuda()
info = [{'speaker': {'index': i}} for i in range(3)]
generate._GREEDY = True
ya, la = model.test(utt, info, return_logp=True, use_graphs=False)
yb, lb = model.test(utt, info, return_logp=True, use_graphs=True)
fs = 12
print('samples equal:', torch.equal(ya, yb))
d = (la - lb).abs().amax(dim=(0, 2))
for f in range(5):
    print('frame', f, ['%.3f' % float(d[f * fs + p]) for p in range(fs)])
print('first diff sample idx', (ya != yb).any(0).nonzero()[:3].flatten().tolist())
