import torch, os, sys
sys.path.insert(0,'/root/repo')
from ta3n_b200 import _lib
import ta3n_b200
lib=_lib.load(); ta3n_b200.set_gemm_engine("tf32")
dev=torch.device("cuda:0")
g=torch.Generator().manual_seed(1)
for (M,N,K) in [(512,256,2560),(2560,512,2048)]:
    A=torch.randn(M,K,generator=g).abs().to(dev); B=(torch.randn(N,K,generator=g)*0.02).to(dev)
    C=torch.empty(M,N,device=dev)
    _lib.check(lib.ta3n_gemm_tn(A.data_ptr(),B.data_ptr(),C.data_ptr(),M,N,K,torch.cuda.current_stream().cuda_stream))
    ref=(A.double()@B.double().t())
    e=((C.double()-ref).norm()/ref.norm()).item()
    bias=((C.double()-ref)*ref).sum().item()/(ref*ref).sum().item()
    print(os.environ.get("TA3N_TMA_RAW_FP32","0"),M,N,K,"normwise",e,"relative bias",bias)
