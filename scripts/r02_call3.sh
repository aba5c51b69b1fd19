#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r02_pytest_3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_3.log
tail -n 15 gpurun_out/r02_pytest_3.log
{ nvidia-smi topo -m; lscpu | head -40; ls /sys/devices/system/node/; cat /sys/devices/system/node/node*/meminfo | grep -i "MemTotal\|MemFree"; nproc; cat /proc/self/status | grep -i "cpus_allowed_list\|mems_allowed_list"; numactl -H 2>&1 | head; free -g; nvidia-smi -q | grep -i -A3 "pci" | head -60; } > gpurun_out/r02_host_topology.txt 2>&1
