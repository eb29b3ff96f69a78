set -x
nvidia-smi topo -m
nvidia-smi -q | grep -i -B1 -A4 -E "encoder|decoder|jpeg|ofa" | head -80
nvidia-smi --query-gpu=index,pci.bus_id,name --format=csv
lscpu | head -40
nproc
cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null
cat /sys/fs/cgroup/cpu.max 2>/dev/null
numactl -H 2>&1 | head -20
ls /sys/devices/system/node/ 2>/dev/null
for n in /sys/devices/system/node/node*; do echo $n $(cat $n/cpulist) $(grep MemTotal $n/meminfo); done
for d in /sys/bus/pci/devices/*; do v=$(cat $d/vendor); c=$(cat $d/class); if [ "$v" = "0x10de" ]; then echo $d $c numa=$(cat $d/numa_node) local_cpus=$(cat $d/local_cpulist); fi; done
free -g
ulimit -l
python -c "import os; print(os.sched_getaffinity(0))"
ls /usr/lib/x86_64-linux-gnu | grep -i -E "nvcuvid|nvidia-encode|nvjpeg|numa" 
ls /usr/local/cuda/lib64 | grep -i -E "nvjpeg|npp" | head
