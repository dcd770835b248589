"""cProfile of the host side of one ISW module step (InstanceWhitening + loss, forward + backward)."""
import cProfile, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dgvcc_b200.models.ISW import InstanceWhitening, instance_whitening_loss

dev = torch.device("cuda:0")
b, c, h, w = 8, 256, 80, 80
xin = torch.randn(b, c, h, w, device=dev, requires_grad=True)
eye = torch.eye(c, device=dev)
mask = torch.triu(torch.ones(c, c, device=dev), 1)
num = mask.sum()
iw = InstanceWhitening(c)


def step():
    xin.grad = None
    y, wt = iw(xin)
    loss = instance_whitening_loss(wt, eye, mask, 0, num)
    loss.backward()


for _ in range(20):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)
