import torch, numpy as np
x = torch.linspace(3, 10, 200001, dtype=torch.float32)
ref = torch.tanh(x.double())
cpu = torch.tanh(x)
gpu = torch.tanh(x.cuda()).cpu()
def stat(name, a):
    la = torch.log(1 - a.double()**2 + 1e-6) if False else torch.log((1 - a*a + 1e-6).double())
    lr = torch.log(1 - ref**2 + 1e-6)
    d = (la - lr)
    ulp = (a.double() - ref) / 5.96e-8
    print(name, 'mean dlog %.4e' % d.mean().item(), 'max|dlog| %.3f' % d.abs().max().item(), 'mean ulp err %.3f' % ulp.mean().item(), 'max ulp %.2f' % ulp.abs().max().item(), 'frac==1: %.3f' % (a == 1).float().mean().item())
stat('cpu', cpu); stat('gpu', gpu)
cr = ref.float()
stat('rounded-f64', cr)
