#!/bin/bash
# usage: tools/ptxinfo.sh file.cu [extra nvcc flags] -> per kernel: registers, spills, shared memory (ptxas -v, sm_100a)
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -Xptxas=-v "$@" -c -o /tmp/ptxinfo.o 2>&1 | python3 -c "
import sys,re,subprocess
name=None; spill=''
for l in sys.stdin:
    m=re.search(r\"Compiling entry function '(\S+)'\",l)
    if m: name=m.group(1)
    m=re.search(r'(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads',l)
    if m: spill='stack %s spill %s/%s'%m.groups()
    m=re.search(r'Used (\d+) registers(.*)',l)
    if m and name:
        d=subprocess.run(['c++filt',name],capture_output=True,text=True).stdout.strip()
        print('%-90s regs %s %s %s'%(d[:90],m.group(1),spill,m.group(2).strip()[:60])); name=None
    if 'error' in l or 'warning' in l: print(l.rstrip())
"
