import os, sys, ctypes
os.environ["PSB_DEBUG_TIMING"]="1"
sys.path.insert(0,"/root/repo")
import numpy as np, torch
import bench
dev=torch.device("cuda:0"); stream=torch.cuda.Stream()
for wl,natoms in (("abf2d",100_000),("harmonic",10_000),("harmonic",1_000_000)):
    frames,L,_=bench.device_hoomd_frames(natoms,8,dev,seed=1)
    snaps=bench.snapshots_of(frames,L,0.005)
    cvs,m=bench.workload_specs(wl,natoms,L)
    method,native=bench.build_ours(cvs,m,snaps[0])
    descs=bench.snapdesc_array(native,snaps)
    ms,_=bench.time_cycle(native,descs,200,20,stream)
    torch.cuda.synchronize()
    t=native.view("debug_timing",(-1,16)).cpu().numpy()
    g=t.shape[0]
    t0=t[:,0].min()
    fin=[b for b in range(g) if t[b,8]>=t0]  # finalizing block in last launch (others hold stale stamps)
    fb=max(range(g), key=lambda b:t[b,8])
    rel=lambda x:(x-t0)/1e3
    print(f"{wl} N={natoms} grid={g} us/step={ms*1e3/200:.2f}")
    print("  start spread   : %.2f us"%rel(t[:,0].max()))
    print("  passA done     : median %.2f max %.2f"%(np.median(rel(t[:,1])),rel(t[:,1].max())))
    print("  finalize blk %d : enter %.2f reduce %.2f cvsA %.2f postcv %.2f store %.2f"%(fb,*[rel(t[fb,i]) for i in (8,9,10,11,12)]))
    print("  released       : median %.2f max %.2f"%(np.median(rel(t[:,2])),rel(t[:,2].max())))
    print("  scatter start  : max %.2f ; end median %.2f max %.2f"%(rel(t[:,3].max()),np.median(rel(t[:,4])),rel(t[:,4].max())))
