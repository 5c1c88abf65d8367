#!/bin/bash
# Environment probe: dumps everything the design branches on.
out=gpurun_out/probe_env.txt
mkdir -p gpurun_out
{
echo "=== date"; date
echo "=== id"; id; grep -E 'Cap(Eff|Prm|Bnd)' /proc/self/status
echo "=== uname"; uname -a
echo "=== nvidia-smi -L"; nvidia-smi -L
echo "=== nvidia-smi topo -m"; nvidia-smi topo -m
echo "=== nvidia-smi nvlink -s (gpu0)"; nvidia-smi nvlink -s -i 0 | head -30
echo "=== /dev/infiniband"; ls -la /dev/infiniband 2>&1
echo "=== /sys/class/infiniband"; ls -la /sys/class/infiniband 2>&1
echo "=== /sys/class/infiniband_verbs"; ls -la /sys/class/infiniband_verbs 2>&1
echo "=== /sys/class/net"; ls /sys/class/net 2>&1
echo "=== PCI mellanox"; for d in /sys/bus/pci/devices/*; do v=$(cat $d/vendor 2>/dev/null); if [ "$v" = "0x15b3" ]; then echo "$d $(cat $d/device) $(cat $d/class)"; fi; done
echo "=== PCI nvidia"; for d in /sys/bus/pci/devices/*; do v=$(cat $d/vendor 2>/dev/null); if [ "$v" = "0x10de" ]; then echo "$d $(cat $d/device) $(cat $d/class) numa=$(cat $d/numa_node 2>/dev/null)"; fi; done
echo "=== /proc/modules"; grep -E 'nvidia|mlx5|ib_|peermem|gdrdrv|rdma' /proc/modules 2>&1
echo "=== peermem"; cat /sys/module/nvidia_peermem/version 2>&1; ls -R /sys/kernel/mm/memory_peers 2>&1
echo "=== ldconfig"; ldconfig -p | grep -E 'ibverbs|mlx5|rdmacm|nccl|cuda|gdr' 
echo "=== find libs"; find / -xdev \( -name 'libibverbs*' -o -name 'libmlx5*' -o -name 'librdmacm*' -o -name 'libgdrapi*' -o -name 'nv-p2p.h' -o -name 'peer_mem.h' \) 2>/dev/null | head -40
echo "=== nvidia params"; cat /proc/driver/nvidia/params 2>&1 | head -80
echo "=== nvidia version"; cat /proc/driver/nvidia/version 2>&1
echo "=== /lib/modules"; ls /lib/modules 2>&1
echo "=== iommu groups"; ls /sys/kernel/iommu_groups 2>&1 | wc -l
echo "=== /dev nvidia"; ls -la /dev | grep -iE 'nvidia|infiniband|gdr|dma_heap|vfio|fuse'
echo "=== /dev/nvidia-caps"; ls -la /dev/nvidia-caps 2>&1
echo "=== nvidia-smi -q (gpu0 head)"; nvidia-smi -q -i 0 | head -150
echo "=== nvidia-smi clocks"; nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,pci.bus_id,memory.total --format=csv
echo "=== cpu/mem"; nproc; free -g; lscpu | head -25
echo "=== mounts"; cat /proc/mounts | head -60
echo "=== ulimit"; ulimit -a
echo "=== env"; env | grep -iE 'nccl|cuda|nvidia|graft' 
echo "=== shm"; df -h /dev/shm /tmp
echo "=== nvcc"; nvcc --version | tail -2
} > $out 2>&1
python - <<'PY' >> gpurun_out/probe_env.txt 2>&1
import torch, time
print("=== torch", torch.__version__, torch.cuda.is_available(), torch.cuda.device_count())
for i in range(torch.cuda.device_count()):
    p = torch.cuda.get_device_properties(i)
    print(i, p.name, p.total_memory, p.multi_processor_count, p.major, p.minor)
if torch.cuda.device_count() > 1:
    print("p2p 0->1", torch.cuda.can_device_access_peer(0,1))
from cuda.bindings import driver as cu
print(cu.cuInit(0))
err, dev = cu.cuDeviceGet(0)
for name in ["CU_DEVICE_ATTRIBUTE_DMA_BUF_SUPPORTED","CU_DEVICE_ATTRIBUTE_GPU_DIRECT_RDMA_SUPPORTED","CU_DEVICE_ATTRIBUTE_GPU_DIRECT_RDMA_FLUSH_WRITES_OPTIONS","CU_DEVICE_ATTRIBUTE_GPU_DIRECT_RDMA_WRITES_ORDERING","CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED","CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED","CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED","CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED","CU_DEVICE_ATTRIBUTE_CAN_USE_HOST_POINTER_FOR_REGISTERED_MEM","CU_DEVICE_ATTRIBUTE_HOST_REGISTER_SUPPORTED","CU_DEVICE_ATTRIBUTE_CAN_USE_STREAM_WAIT_VALUE_NOR","CU_DEVICE_ATTRIBUTE_MEM_SYNC_DOMAIN_COUNT","CU_DEVICE_ATTRIBUTE_PAGEABLE_MEMORY_ACCESS","CU_DEVICE_ATTRIBUTE_HOST_NATIVE_ATOMIC_SUPPORTED","CU_DEVICE_ATTRIBUTE_MAX_SHARED_MEMORY_PER_BLOCK_OPTIN","CU_DEVICE_ATTRIBUTE_L2_CACHE_SIZE","CU_DEVICE_ATTRIBUTE_CLUSTER_LAUNCH"]:
    a = getattr(cu.CUdevice_attribute, name, None)
    if a is None: print(name, "n/a"); continue
    print(name, cu.cuDeviceGetAttribute(a, dev))
PY
echo done
