# what the host of a GPU box looks like (NUMA placement of the GPUs, memory policy of the container)
nvidia-smi topo -m 2>&1 | head -14
lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)"
grep -E "Cpus_allowed_list|Mems_allowed_list" /proc/self/status
ls /sys/devices/system/node/ | head
for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor)" = "0x10de" ]; then echo $d $(cat $d/class) numa=$(cat $d/numa_node) local_cpus=$(cat $d/local_cpulist); fi; done 2>/dev/null | head -20
free -g | head -2
