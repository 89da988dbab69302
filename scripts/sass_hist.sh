#!/bin/bash
# SASS opcode histogram of the shipped library (whole library, then the kernels the bench times) -> profiles/r02_sass_histogram.txt
cd "$(dirname "$0")/.."
SO=mimsem_b200/libmimsem_gpu.so
OUT=profiles/r02_sass_histogram.txt
hist() { grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+(@!?U?P[0-9T]+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]*)?.*/\2/' | sort | uniq -c | sort -rn; }
{
  echo "# cuobjdump -sass $SO ($(date -u +%FT%TZ)); nvcc $(nvcc --version | grep release | sed 's/.*release //')"
  echo "## whole library: opcodes that identify the async-copy / barrier / FP64 paths"
  cuobjdump -sass $SO | hist | grep -E " (UBLKCP|UBLKPF|SYNCS|DFMA|DMUL|DADD|UTMALDG|UTCMMA|LDGSTS|ACQBULK|ERRBAR|MEMBAR|FENCE|LDG|STG|LDS|STS|RED|ATOM|ATOMG|BAR|ELECT|SETMAXREG)$"
  for k in "k_apply_m1_pipeILi4ELb0ELi60ELi1EE" "k_apply_m1_tileILi4ELb0ELi60ELi0ELi4ELi1EE" "k_apply_m1_tileILi4ELb1ELi60ELi0ELi4ELi2EE" "k_apply_m1_tileILi4ELb0ELi60ELi1ELi4ELi1EE" "k_apply_k_tmaILi4ELi60EE" "k_apply_m2_tileILi4ELb0ELi60EE"; do
    f=$(cuobjdump -elf $SO 2>/dev/null | grep -o "_ZN6mimsem[0-9]*${k}[A-Za-z0-9_]*" | sort -u | head -1)
    [ -z "$f" ] && continue
    echo "## $(echo $f | c++filt)"
    cuobjdump -sass -fun "$f" $SO 2>/dev/null | hist | head -24
  done
} > $OUT
wc -l $OUT; head -40 $OUT
