import sys, numpy as np, torch, time
sys.path.insert(0,'.')
from quantum_inferno_b200 import cwt_entropy
sys.path.insert(0,'tests')
from tests.test_gpu_parity import synth, FS
for order, logn, C in [(3,20,2),(6,18,2),(12,16,1),(3,24,1)]:
    n=1<<logn
    x=torch.from_numpy(np.stack([synth(n,chan=c) for c in range(C)])).cuda()
    a=cwt_entropy.cwt_power_entropy(order,x,FS,dtype="float32",method="multirate")
    pa=a.power.double().clone(); ea=a.entropy_bits().clone(); ta=a.total_power.clone()
    del a
    b=cwt_entropy.cwt_power_entropy(order,x,FS,dtype="float32",method="exact")
    pb=b.power.double()
    l2=float((pa-pb).norm()/pb.norm())
    perband=((pa-pb).norm(dim=2)/pb.norm(dim=2)).max(dim=0).values
    print(f"order {order} 2^{logn} C={C}: B={pa.shape[1]} multirate-vs-exact power L2 {l2:.2e}; entropy diff {float((ea-b.entropy_bits()).abs().max()):.2e} bits; total rel {float(((ta-b.total_power)/b.total_power).abs().max()):.2e}; per-band L2 max {float(perband.max()):.2e} (band {int(perband.argmax())}), excluding 4 lowest {float(perband[4:].max()):.2e}")
    del pa,pb,b; torch.cuda.empty_cache()
