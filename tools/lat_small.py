import sys, time, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import u96_slam_b200 as u
from oracle_py import Oracle
o=Oracle()
for (W,H,D,B,n) in ((640,480,64,21,1),(640,480,64,21,4),(640,480,64,21,8),(1242,375,128,21,1),(640,480,64,15,1)):
    L,R=u.synth_batch(3,0,n,W,H,D)
    with u.StereoFrontEnd(0,W,H,n) as fe:
        fe.set_bm_params(width=W,height=H,profile=0,block_size=B,num_disparities=D,x_store_offset=1,uni_enable=0)
        fe.set_profiling(True)
        for i in range(5):
            fe.submit_rect(i&1,L,R); b=fe.wait()
        st=fe.last_stage_ms_ex(b)
        d=fe.receive_disp(b); xl,xr=fe.receive_xsbl(b)
        bad=sum(int((d[i]!=o.bm_rtl(xl[i],xr[i],wsz=B,ndisp=D)).sum()) for i in range(n))
        fe.set_profiling(False)
        t=[]
        for i in range(30):
            t0=time.perf_counter(); fe.submit_rect(i&1,L,R); b=fe.wait(); dd=fe.receive_disp(b); t.append((time.perf_counter()-t0)*1e3)
        print(W,H,D,B,'n',n,'bm_ms %.4f'%st['bm'],'mismatches',bad,'e2e_ms median %.3f'%np.median(t), flush=True)
