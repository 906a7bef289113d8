#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
nproc; lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)"
nvidia-smi topo -m 2>&1 | head -20
python - <<'PY'
import torch, os
p=torch.cuda.get_device_properties(0)
print([a for a in dir(p) if 'pci' in a.lower()])
for a in ('pci_bus_id','pci_device_id','pci_domain_id'):
    print(a, getattr(p,a,None))
try:
    import pynvml
    pynvml.nvmlInit()
    for i in range(pynvml.nvmlDeviceGetCount()):
        h=pynvml.nvmlDeviceGetHandleByIndex(i)
        b=pynvml.nvmlDeviceGetPciInfo(h).busId
        b=b.decode() if isinstance(b,bytes) else b
        path='/sys/bus/pci/devices/'+b[-12:].lower()+'/numa_node'
        try: nn=open(path).read().strip()
        except Exception as e: nn=repr(e)
        print(i,b,path,nn)
except Exception as e: print('nvml',e)
print(len(os.sched_getaffinity(0)))
for n in range(4):
    try: print(n, open(f'/sys/devices/system/node/node{n}/cpulist').read().strip())
    except Exception as e: break
PY
} > gpurun_out/topo.txt 2>&1
cat gpurun_out/topo.txt
