import os, sys, numpy as np
os.environ["GPUCHAN_DEBUG_STAMPS"]="1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))); import tslb200_loader; tslb200_loader.load_package()
from tsl_sdr_b200 import synth
from tsl_sdr_b200.gpuchan import GpuChan
import ctypes as C
fs,T,D,Cn=2400000,127,100,64
n=1<<24
lpf=synth.lowpass_taps(T,9000.0,fs); offs=synth.channel_offsets(Cn,fs)
rng=np.random.default_rng(1); iq=np.clip(np.round(rng.normal(0,3000,2*n)),-32768,32767).astype(np.int16)
b=GpuChan(lpf,offs,fs,D,n)
for i in range(3):
    b.submit(iq); b.collect()
out=np.zeros((3,32,8),np.int64)
assert b._L.gpuchan_debug_stamps(b._h, out.ctypes.data)==0
t0=out[out>0].min()
o=np.where(out>0,out-t0,-1)
np.set_printoptions(linewidth=200)
print("producer [start, after b_empty, after copies]"); print(o[0,:12,:3])
print("mma [start, after b_full, after t_empty, after issue, after commit]"); print(o[1,:12,:5])
print("epilogue [start, after t_full, after phase1, after bar1, after phase2, after bar2]"); print(o[2,:12,:6])
