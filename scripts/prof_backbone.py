import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eas_snn_b200 import fused
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = fused.SpikingCSPDarknet(0.67, 0.75, in_dim=2, T=3).to(dev).eval()
for m in net.modules():
    if isinstance(m, torch.nn.BatchNorm2d):
        m.bias.data.fill_(0.6)
x = torch.rand(1, 64, 2, 256, 320, device=dev) * 2
for _ in range(2):
    net(x)
torch.cuda.synchronize()
