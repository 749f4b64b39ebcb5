import sys, os, ctypes
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import numpy as np, torch
from synth import *
import depthmap_b200._lib as L
from depthmap_b200.stereoimage_generation import create_stereoimages
from oracle import stereo as ost
from reformulation_model import _vertex_x
lib = L.load()
lib.dm_stereo_set_debug_buffer.argtypes = [ctypes.c_void_p]
h, w = 2, 8
img = synth_rgb(h, w, 0); dep = synth_depth_u16(h, w, 0)
for fill in ['polylines_soft', 'polylines_sharp']:
    dbg = torch.zeros(4096, dtype=torch.float64, device='cuda')
    lib.dm_stereo_set_debug_buffer(dbg.data_ptr())
    got = np.asarray(create_stereoimages(img, dep, 0.0001, 0.0, ['left-right'], 0.0, 1.0, fill)[0])
    torch.cuda.synchronize()
    lib.dm_stereo_set_debug_buffer(None)
    d = dbg.cpu().numpy(); n = int(d[0])
    print(fill, 'n', n)
    print(' pm   ', np.round(d[1:1+n], 3).tolist())
    print(' order', d[1+n:1+2*n].astype(int).tolist())
    print(' X    ', np.round(d[1+2*n:1+3*n], 3).tolist())
    print(' off  ', d[1+3*n:1+3*n+w+2].astype(int).tolist())
    want = ost.create_stereoimages(img, dep, 0.0001, 0.0, ['left-right'], 0.0, 1.0, fill, return_arrays=True)[0]
    print(' got ', got[0, :w, 0].tolist()); print(' want', want[0, :w, 0].tolist())
