import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch, numpy as np
from oracle import fsnet_oracle as O
from helpers import build_model
from test_oracle_golden import FULL_CASES, rel
torch.backends.cudnn.allow_tf32=False
for name in ['tiny_r50','tiny_sigmoid','tiny4']:
    g=np.load(f'/root/repo/tests/golden/{name}.npz'); topo=FULL_CASES[name]['topo']
    data=O.synthetic_batch(2,topo.height,topo.width,1234,topo.frame_ids)
    m=build_model(topo).cuda()
    feats=m.depth_backbone(data[('image',0)].cuda())
    print(name,[ (float(f.abs().mean()), float(g[f'feat_absmean/{i}'])) for i,f in enumerate(feats)])
    outs=m.head.forward_depth(feats, data['P2'].cuda())
    print([rel(outs[('disp',s)].detach().cpu(), g[f'disp/{s}']) for s in topo.scales])
    # cpu fp32 same modules
    sd=O.make_state_dict(topo)
    with torch.no_grad():
        f2=O.resnet_forward(sd,'depth_backbone.',data[('image',0)],topo.depth)
    print('feat rel gpu vs cpu oracle',[rel(a.detach().cpu(),b) for a,b in zip(feats,f2)])
