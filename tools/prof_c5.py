import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
from tools.measure_configs import load_b2nn
b2nn = load_b2nn()
net = b2nn.Net(0)
net.set_layers([784, 50, 10], b2nn.INIT_2, 42)
d = b2nn.make_desc(classify=True, act=b2nn.ACT_TANH, out=b2nn.OUT_SOFTMAX, precision=b2nn.PREC_TF32)
Bn = 1 << 18
xd = torch.randn(785, Bn, device="cuda"); xd[784] = 1.0
outd = torch.empty(Bn, device="cuda")
for _ in range(4):
    net.predict_device(d, xd.data_ptr(), Bn, Bn, outd.data_ptr())
torch.cuda.synchronize()
print("ok")
