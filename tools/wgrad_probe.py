import torch, sys
sys.path.insert(0,'/root/repo')
from ta3n_b200 import _lib
import ta3n_b200
lib=_lib.load(); ta3n_b200.set_gemm_engine("tf32")
dev=torch.device("cuda:0"); st=torch.cuda.current_stream().cuda_stream
g=torch.Generator().manual_seed(1)
def gemm_ex(A,lda,ak,B,ldb,bk,M,N,K,ws=None):
    C=torch.zeros(M,N,device=dev)
    _lib.check(lib.ta3n_gemm_ex(A.data_ptr(),lda,ak,B.data_ptr(),ldb,bk,C.data_ptr(),N,M,N,K, ws.data_ptr() if ws is not None else None, ws.numel() if ws is not None else 0, st))
    torch.cuda.synchronize(); return C
print("wgrad layout (A M-major, B N-major) with K tails")
for (M,N,K) in [(128,128,32),(128,128,40),(128,128,8),(128,128,33),(512,2048,80),(512,2048,185),(256,256,100)]:
    At=torch.randn(K,M,generator=g).to(dev); B=torch.randn(K,N,generator=g).to(dev)
    C=gemm_ex(At,M,0,B,N,0,M,N,K)
    ref=At.double().t()@B.double()
    err=(C.double()-ref)
    print(M,N,K,"rel",(err.norm()/ref.norm()).item(), "max row-block err", [round((err[i:i+32].norm()/ref[i:i+32].norm()).item(),5) for i in range(0,min(M,128),32)])
print("dgrad layout (A K-major, B N-major) with K tails")
for (M,N,K) in [(128,128,40),(256,512,100)]:
    A=torch.randn(M,K,generator=g).to(dev); B=torch.randn(K,N,generator=g).to(dev)
    C=gemm_ex(A,K,1,B,N,0,M,N,K)
    ref=A.double()@B.double()
    print(M,N,K,"rel",((C.double()-ref).norm()/ref.norm()).item())
print("fwd layout K tails")
for (M,N,K) in [(128,128,40),(256,512,100)]:
    A=torch.randn(M,K,generator=g).to(dev); B=torch.randn(N,K,generator=g).to(dev)
    C=gemm_ex(A,K,1,B,K,1,M,N,K)
    ref=A.double()@B.double().t()
    print(M,N,K,"rel",((C.double()-ref).norm()/ref.norm()).item())
