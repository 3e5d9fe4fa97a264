import importlib, sys, torch
sys.path.insert(0, '/root/repo')
tr = importlib.import_module("3dal_pytorch_b200.train")
torch.manual_seed(0)
for M in (5120, 4096, 404, 20, 4, 5000):
    for C in (64, 128, 256, 512, 1024):
        for off in (0.5, 30.0):
            y = (torch.randn(M, C, device="cuda") * 2 + off * torch.randn(1, C, device="cuda")).contiguous()
            bn = torch.nn.BatchNorm1d(C).cuda()
            z, mean, rstd = tr.bn_forward(y, bn, relu=True)
            m64 = y.double().mean(0); v64 = y.double().var(0, unbiased=False)
            r64 = 1.0 / torch.sqrt(v64 + bn.eps)
            em = float((mean.double() - m64).abs().max() / m64.abs().max()); er = float((rstd.double() - r64).abs().max() / r64.abs().max())
            flag = "  <<<<" if em > 1e-5 or er > 1e-4 else ""
            print("M %5d C %4d off %4.1f  mean err %.2e  rstd err %.2e%s" % (M, C, off, em, er, flag))
