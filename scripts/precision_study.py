import sys, importlib, torch, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from helpers import fold_state_dict, synth, spec
from oracle import models
torch.set_num_threads(8)
def r_bf16(t): return t.to(torch.bfloat16).float()
def r_f16(t): return t.half().float()
def r_mant(bits):
    def f(t):
        # round-to-nearest-even to `bits` explicit mantissa bits
        i = t.contiguous().view(torch.int32)
        drop = 23-bits
        bias = ((i >> drop) & 1) + (1 << (drop-1)) - 1
        return (((i + bias) >> drop) << drop).view(torch.float32)
    return f
r_tf32 = r_mant(10)
ident = lambda t: t
def seg(fw, x, ra, rw, ra_last=None):
    h = x.transpose(2,1).float()
    h = ra(torch.relu(h @ fw['conv1'][0].t() + fw['conv1'][1]))
    o2 = ra(torch.relu(h @ rw(fw['conv2'][0]).t() + fw['conv2'][1]))
    o3 = ra(torch.relu(o2 @ rw(fw['conv3'][0]).t() + fw['conv3'][1]))
    o4 = ra(torch.relu(o3 @ rw(fw['conv4'][0]).t() + fw['conv4'][1]))
    g = torch.relu((o4 @ rw(fw['conv5'][0]).t()).max(dim=1)[0] + fw['conv5'][1])
    wd1,bd1 = fw['dconv1']
    gb = g @ wd1[:,64:].t() + bd1
    d = ra(torch.relu(o2 @ rw(wd1[:,:64]).t() + gb[:,None,:]))
    d = ra(torch.relu(d @ rw(fw['dconv2'][0]).t() + fw['dconv2'][1]))
    d = ra(torch.relu(d @ rw(fw['dconv3'][0]).t() + fw['dconv3'][1]))
    d = torch.relu(d @ rw(fw['dconv4'][0]).t() + fw['dconv4'][1])
    return d @ fw['dconv5'][0].t() + fw['dconv5'][1], g
for seed in (synth.REFERENCE_SEED, 3):
    sd = synth.random_state_dict('static_one', seed=seed)
    d = synth.static_tracks(32, n=4096, seed=7)
    pts = torch.from_numpy(d['pts_pm']).transpose(2,1)
    lg,_ = models.seg_forward(sd, pts)
    synth.calibrate_seg_margin(sd, lg, 0.3)
    fw = fold_state_dict(sd, 'ins_seg', spec.seg_layers(3))
    ref, gref = seg(fw, pts.double() if False else pts, ident, ident)
    lg2,_ = models.seg_forward(sd, pts)
    print('seed',seed,'folded-vs-oracle', float((ref-lg2).abs().max()/lg2.abs().max()), 'max|logit|', float(lg2.abs().max()), 'margin std', float((lg2[...,1]-lg2[...,0]).std()))
    m_ref = lg2[...,0] < lg2[...,1]
    for name, ra, rw in [('bf16', r_bf16, r_bf16), ('fp16', r_f16, r_f16), ('tf32', r_tf32, r_tf32),
                         ('fp16 act, fp32 w', r_f16, ident), ('fp32 act, fp16 w', ident, r_f16),
                         ('m13', r_mant(13), r_mant(13)), ('m16 (bf16 hi+lo)', r_mant(16), r_mant(16))]:
        l, g = seg(fw, pts, ra, rw)
        m = l[...,0] < l[...,1]
        fl = (m != m_ref)
        marg = (lg2[...,1]-lg2[...,0])
        print('  %-20s logits rel %.2e  g rel %.2e  flips %d / %d  max|margin| flipped %.2e' % (name, float((l-lg2).abs().max()/lg2.abs().max()), float((g-gref).abs().max()/gref.abs().max()), int(fl.sum()), fl.numel(), float(marg[fl].abs().max()) if fl.any() else 0))

print("---- per-layer sensitivity (only that layer's A-operand + weights rounded)")
def seg_sel(fw, x, sel, r):
    # sel: set of layer names whose input activations and weights are rounded by r
    def R(name, t): return r(t) if name in sel else t
    h = x.transpose(2,1).float()
    o1 = torch.relu(h @ fw['conv1'][0].t() + fw['conv1'][1])
    o2 = torch.relu(R('conv2',o1) @ R('conv2',fw['conv2'][0]).t() + fw['conv2'][1])
    o3 = torch.relu(R('conv3',o2) @ R('conv3',fw['conv3'][0]).t() + fw['conv3'][1])
    o4 = torch.relu(R('conv4',o3) @ R('conv4',fw['conv4'][0]).t() + fw['conv4'][1])
    g = torch.relu((R('conv5',o4) @ R('conv5',fw['conv5'][0]).t()).max(dim=1)[0] + fw['conv5'][1])
    wd1,bd1 = fw['dconv1']
    gb = g @ wd1[:,64:].t() + bd1
    d = torch.relu(R('dconv1',o2) @ R('dconv1',wd1[:,:64]).t() + gb[:,None,:])
    d = torch.relu(R('dconv2',d) @ R('dconv2',fw['dconv2'][0]).t() + fw['dconv2'][1])
    d = torch.relu(R('dconv3',d) @ R('dconv3',fw['dconv3'][0]).t() + fw['dconv3'][1])
    d = torch.relu(R('dconv4',d) @ R('dconv4',fw['dconv4'][0]).t() + fw['dconv4'][1])
    return d @ fw['dconv5'][0].t() + fw['dconv5'][1], g
sd = synth.random_state_dict('static_one', seed=synth.REFERENCE_SEED)
d = synth.static_tracks(32, n=4096, seed=7)
pts = torch.from_numpy(d['pts_pm']).transpose(2,1)
lg,_ = models.seg_forward(sd, pts)
synth.calibrate_seg_margin(sd, lg, 0.3)
fw = fold_state_dict(sd, 'ins_seg', spec.seg_layers(3))
lg2,_ = models.seg_forward(sd, pts)
names=['conv2','conv3','conv4','conv5','dconv1','dconv2','dconv3','dconv4']
for rn, r in (('bf16', r_bf16), ('fp16', r_f16)):
    for nm in names:
        l,g = seg_sel(fw, pts, {nm}, r)
        print('  %s only %-7s logits rel %.2e' % (rn, nm, float((l-lg2).abs().max()/lg2.abs().max())))
    for keep in (['conv5'], ['conv5','dconv2'], ['conv2','conv3','conv4','conv5'], ['conv5','dconv1','dconv2']):
        l,g = seg_sel(fw, pts, set(keep), r)
        print('  %s only %s logits rel %.2e' % (rn, keep, float((l-lg2).abs().max()/lg2.abs().max())))
