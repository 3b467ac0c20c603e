import ctypes as C, numpy as np, sys
sys.path.insert(0,'/root/repo')
import mlvfs_b200 as M
from mlvfs_b200 import synth, mlvformat as F
from oracle import pyoracle as O
w,h=1920,1080
L=M.lib(); L.cr2hdr20_convert_data.argtypes=[C.c_void_p,C.c_void_p,C.c_int,C.c_int,C.c_int,C.c_int,C.c_int]
img=synth.make_frame(w,h,0,dual_iso=True,hot_cold=True,bad_density=1e-4)
for (cs,alias) in [(0,0),(0,1),(5,0),(3,0),(2,0),(5,1)]:
    hdr=F.make_frame_headers(w,h,file_guid=0xB000+cs*4+alias)
    rc,want,info=O.cr2hdr20(img,2048,15000,interp_method=1,use_alias_map=alias,chroma_smooth_method=cs)
    got=img.copy(); r=L.cr2hdr20_convert_data(C.byref(hdr),got.ctypes.data_as(C.c_void_p),1,1,alias,cs,0)
    d=np.abs(got.astype(int)-want.astype(int)); ys,xs=np.nonzero(d)
    print('cs',cs,'alias',alias,'ndiff',len(ys),'max',d.max())
    if len(ys):
        print('  y range',ys.min(),ys.max(),'x range',xs.min(),xs.max())
        print('  sample',list(zip(ys[:12],xs[:12])), 'x%2',np.bincount(xs%2),'y%4',np.bincount(ys%4))
