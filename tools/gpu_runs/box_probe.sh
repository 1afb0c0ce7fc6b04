#!/bin/bash
# prints what the host side of the box looks like (cores, NUMA, memory, GPU<->CPU affinity)
echo "== nproc"; nproc; echo "== cpuset"; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null; cat /sys/fs/cgroup/cpuset.mems.effective 2>/dev/null
echo "== taskset"; taskset -p $$ 
echo "== lscpu"; lscpu | head -30
echo "== numa nodes"; ls /sys/devices/system/node/ | head; for n in /sys/devices/system/node/node*; do echo $n $(cat $n/cpulist) ; grep MemTotal $n/meminfo; done
echo "== mem"; free -g; cat /sys/fs/cgroup/memory.max 2>/dev/null
echo "== topo"; nvidia-smi topo -m
echo "== gpus"; nvidia-smi --query-gpu=index,pci.bus_id,name,memory.total --format=csv
for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ] && [ "$(cat $d/class 2>/dev/null | cut -c1-6)" = "0x0302" ]; then echo "$d numa=$(cat $d/numa_node) local_cpus=$(cat $d/local_cpulist)"; fi; done
echo "== numactl"; which numactl; numactl -H 2>/dev/null | head -20
python - <<'PY'
import ctypes, os
libc = ctypes.CDLL(None, use_errno=True)
# get_mempolicy syscall 239 on x86_64
mode = ctypes.c_int(-1)
r = libc.syscall(239, ctypes.byref(mode), None, 0, None, 0)
print("get_mempolicy rc", r, "mode", mode.value, "errno", ctypes.get_errno())
print("sched_getaffinity", sorted(os.sched_getaffinity(0))[:8], "...", len(os.sched_getaffinity(0)))
PY
