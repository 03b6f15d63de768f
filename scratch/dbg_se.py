import torch, sys
sys.path.insert(0,'/root/repo')
import torch.nn.functional as F
from emotiongestures_b200 import TED
from emotiongestures_b200.engine import Engine
eng=Engine(TED,'cuda:0','tc')
for c,h,w in [(32,128,70),(128,32,18)]:
    g=torch.Generator().manual_seed(c+h); b=5
    x=torch.randn(b,c,h,w,generator=g); wt=torch.randn(c,c,3,3,generator=g)/(c*9)**0.5
    scale=torch.rand(c,generator=g)+0.5; shift=torch.randn(c,generator=g)*0.1
    got,sums=eng.debug_conv_tc(x,wt,scale,shift,se_sums=True)
    ref=F.conv2d(x.half().double(),wt.half().double(),padding=1)*scale.double().view(1,-1,1,1)+shift.double().view(1,-1,1,1)
    m=sums.cpu().double()/(h*w); r=ref.mean(dim=(2,3))
    print(c,h,w,'mean err',(m-r).abs().max().item(),'ref mean absmax',r.abs().max().item())
    print(' got',m[0,:6].tolist()); print(' ref',r[0,:6].tolist())
    print(' ratio', (m[0,:6]/r[0,:6]).tolist())
