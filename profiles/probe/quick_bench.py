import sys, json, numpy as np, torch
sys.path.insert(0, '.')
from qiskit_dynamics_b200 import _abi as abi
from oracle import numpy_oracle as orc
def dev(a): return torch.from_numpy(np.ascontiguousarray(a)).cuda()
def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); best=1e30
    for _ in range(reps):
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best=min(best,e0.elapsed_time(e1))
    return best
for (n,K,B,S) in ((128,8,4096,200),(128,8,4736,200),(128,8,512,200),(32,8,1024,500),(64,8,4096,200),(256,4,4096,50)):
    H0,Hs,Y,sig=orc.synthetic_schrodinger(n,K,B,2004)
    Gd,G,d,U=orc.generator_model_operators(H0,Hs,H0)
    specs=[orc.SigSpec(*s) for s in sig]
    h=1e-3; times=orc.stage_time_grid(0.0,h,S); coeff=dev(orc.signal_list_values(specs,times))
    Gdv,Gdd=dev(G),dev(Gd); Gp,Gdp=abi.pack_operators(Gdv),abi.pack_operators(Gdd[None])[0]
    mu=dev(-np.imag(d)); y=dev(U.conj().T@Y)
    ws=torch.empty(abi.workspace_bytes(abi.WS_RK4,n,K,B,S),dtype=torch.uint8,device='cuda')
    ms=timeit(lambda: abi.rk4_steps(n,Gdv,Gdd,Gp,Gdp,coeff,mu,times,h,y,S,workspace=ws))
    flops=S*B*(4*(8*n*n+12*n)+28*n)
    print(json.dumps({"kind":"rk4_shared","n":n,"K":K,"B":B,"S":S,"ms":ms,"us_per_step":ms*1e3/S,"tflops":flops/ms*1e-9,"state_rhs_per_s":4*S*B/ms*1e3}),flush=True)
    # single RHS call
    c=dev(orc.signal_list_values(specs,0.3)); out=torch.empty_like(y); wsr=torch.empty(abi.workspace_bytes(abi.WS_RHS,n,K,B),dtype=torch.uint8,device='cuda')
    ms=timeit(lambda: abi.rhs(n,Gdv,Gdd,c,mu,0.3,y,out=out,workspace=wsr),reps=20)
    print(json.dumps({"kind":"rhs","n":n,"B":B,"ms":ms,"tflops":B*(8*n*n+12*n)/ms*1e-9}),flush=True)
# sweep
for (n,K,B,S) in ((32,8,1024,200),(128,8,4096,20)):
    H0,Hs,Y,sig=orc.synthetic_schrodinger(n,K,B,2002)
    Gd,G,d,U=orc.generator_model_operators(H0,Hs,H0)
    specs=[orc.SigSpec(*s) for s in sig]
    h=1e-3; times=orc.stage_time_grid(0.0,h,S); base=orc.signal_list_values(specs,times)
    amp=0.5+np.arange(B)/B; coeff=dev(base[:,:,None]*amp[None,None,:])
    Gdv,Gdd=dev(G),dev(Gd); Gp,Gdp=abi.pack_operators(Gdv),abi.pack_operators(Gdd[None])[0]
    mu=dev(-np.imag(d)); y=dev(U.conj().T@Y)
    ms=timeit(lambda: abi.rk4_steps(n,Gdv,Gdd,Gp,Gdp,coeff,mu,times,h,y,S,per_col=True))
    flops=S*B*(4*((4*K+8)*n*n+12*n)+28*n)
    print(json.dumps({"kind":"rk4_sweep","n":n,"K":K,"B":B,"S":S,"ms":ms,"us_per_step":ms*1e3/S,"alg_tflops":flops/ms*1e-9,"exec_tflops":S*B*4*(K+1)*8*n*n/ms*1e-9}),flush=True)
# expm 729
n=729; A=torch.randn(n,n,dtype=torch.complex128,device='cuda'); A=(A-A.conj().T)*0.01
ms=timeit(lambda: abi.expm(A,0)); print(json.dumps({"kind":"expm729_s0","ms":ms,"tflops":6*8*n**3/ms*1e-9}))
Bm=torch.randn(n,4096,dtype=torch.complex128,device='cuda'); C=torch.empty_like(Bm)
ms=timeit(lambda: abi.zgemm(A,Bm,out=C)); print(json.dumps({"kind":"zgemm729x4096","ms":ms,"tflops":8*n*n*4096/ms*1e-9}))
A4=torch.randn(4096,4096,dtype=torch.complex128,device='cuda'); C4=torch.empty_like(A4)
ms=timeit(lambda: abi.zgemm(A4,A4,out=C4)); print(json.dumps({"kind":"zgemm4096","ms":ms,"tflops":8*4096**3/ms*1e-9}))
