import numpy as np, torch, sys
sys.path.insert(0,'/root/repo')
from oracle import restate as R
from tests.golden import inputs
from spatialcore_b200 import engine as eng
coords,X=inputs.g0()
graph,_,_=eng.knn_graph(coords,6)
std=eng.zscore_dense(torch.from_numpy(X).cuda())
num,den,lag,_=eng.lag_moran(graph,std.Z,50)
L=eng.lee_gemm(std.Z,lag,50).cpu().numpy()
W=R.build_spatial_weights(coords,6).astype(np.float64)
Z,_,_,_=R.zscore(X)
want=Z.T@(W@Z)
# exact product of the device FP32 operands in FP64
Zd=std.Z[:,:50].double().cpu().numpy(); Ld=lag[:,:50].double().cpu().numpy()
exact=Zd.T@Ld
print('max|got-want|',np.abs(L-want).max(),' max|got-exact(fp32 inputs)|',np.abs(L-exact).max(),' max|exact-want|',np.abs(exact-want).max())
print('ulp of float32 at 40:', np.spacing(np.float32(40)))
