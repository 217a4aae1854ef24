#!/bin/bash
# what the box says about GPU <-> NUMA placement (for bench.py's bind_to_gpu_numa)
nvidia-smi topo -m 2>&1 | head -30
python - <<'PY'
import torch, os
for i in range(torch.cuda.device_count()):
    p = torch.cuda.get_device_properties(i)
    bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
    base = "/sys/bus/pci/devices/" + bdf
    try:
        print(i, bdf, open(base + "/numa_node").read().strip(), open(base + "/local_cpulist").read().strip())
    except Exception as e:
        print(i, bdf, "ERR", e)
print("affinity", len(os.sched_getaffinity(0)), "nproc", os.cpu_count())
PY
lscpu | grep -i "numa\|socket\|model name" | head
