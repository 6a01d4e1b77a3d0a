# host topology of the GPU box (NUMA nodes, allowed CPUs, GPU affinity): context for the e2e scaling numbers
nproc; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null; lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)"
nvidia-smi topo -m 2>&1 | head -30
for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ]; then echo "$d numa $(cat $d/numa_node) class $(cat $d/class)"; fi; done | head -20
free -g | head -3
