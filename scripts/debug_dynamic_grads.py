import importlib, sys, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import test_gpu_train as T
tr = T.tr
for mode in ("f32", "x3"):
    tr.set_gemm_mode(mode)
    sd, pts, aux, gt, labels = T._case("dynamic")
    sd64, pts64, aux64, gt64, lab64 = T._f64(sd, pts, aux, gt, labels)
    ols, _, ograds, _ = T.otrain.dynamic_step(sd, pts, aux, labels)
    _, _, ograds64, _ = T.otrain.dynamic_step(sd64, pts64, aux64, lab64)
    model = T.dm.DynamicModel().to("cuda:0").train(); model.load_state_dict(sd); model.ins_seg.dropout.p = 0.0
    crit = T.losses.DynamicModelLoss()
    out = model(pts.cuda(), aux.cuda(), gt.cuda())
    ls = crit(out, *[t.cuda() for t in labels]); ls["total_loss"].backward(); torch.cuda.synchronize()
    rows = []
    for name, p in model.named_parameters():
        g, r32, r64 = p.grad.detach().cpu().double(), ograds[name].double(), ograds64[name]
        sc = float(r64.abs().max())
        if sc < 1e-6: continue
        rows.append((float((g - r64).abs().max()) / sc, float((r32 - r64).abs().max()) / sc, sc, name))
    rows.sort(reverse=True)
    print(mode, "pts", tuple(pts.shape))
    for r in rows[:8]: print("  err %.4f  floor %.4f  scale %.3e  %s" % r)
