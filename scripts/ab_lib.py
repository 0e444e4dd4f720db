import os, sys, numpy as np
ROOT='/root/repo'; sys.path.insert(0, ROOT)
from dealiiscale_b200 import capi
if len(sys.argv)>1: capi.LIB_PATH=sys.argv[1]
from dealiiscale_b200.micro import MicroSolver
from dealiiscale_b200.prm import BioData
g=np.linspace(-1,1,65); X,Y=np.meshgrid(g,g,indexing='ij'); xy=np.stack([X.ravel(),Y.ravel()],axis=1)
u,w=xy[:,1]*np.sin(xy[:,0]), xy[:,0]*np.cos(xy[:,1])
ms=MicroSolver(capi.BIO, BioData(os.path.join(ROOT,'input','biocase.prm')),64)
ms.set_grid_locations(xy); ms.setup(); ms.enable_timing(True); ms.set_macro_solution(u,w)
for _ in range(2): ms.zero_solutions(); ms.run()
ms.stats(reset=True)
for _ in range(5): ms.zero_solutions(); its=ms.run()
st=ms.stats(); print(capi.LIB_PATH.split('/')[-1], 'cg %.3f ms'%(st['ms_cg']/5), 'its %.2f'%its.mean())
