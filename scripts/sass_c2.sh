#!/bin/bash
# scripts/sass_c2.sh [extra nvcc flags]: compile only the C2 instantiation (QA_FAST_BUILD) and print where the MUFUs sit
# in the softmax main loop (developer tool: checks the instruction interleave ptxas produced)
cd "$(dirname "$0")/.."
out=/tmp/sass_c2; mkdir -p $out
nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -DQA_FAST_BUILD "$@" -cubin -o $out/a.cubin quantumattention_b200/csrc/attn_fwd.cu 2>&1 | grep -E "error|spill" 
cuobjdump -sass $out/a.cubin > $out/a.sass
python3 - <<'PY'
import re
ins=[]
for l in open('/tmp/sass_c2/a.sass'):
    m=re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);',l)
    if m: ins.append((int(m.group(1),16),re.sub(r'^@!?U?P\d+\s+','',m.group(2).strip())))
# main loop = longest backward branch region among those with >=90 MUFU
best=None
for a,t in ins:
    m=re.search(r'BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)',t)
    if m:
        tgt=int(m.group(1),16)
        if tgt<a:
            body=[x for x in ins if tgt<=x[0]<=a]
            n=sum(1 for x in body if x[1].startswith('MUFU.EX2'))
            if 80<=n<=110 and (best is None or len(body)<len(best)): best=body
print("main loop instrs",len(best))
line=''.join('M' if t.startswith('MUFU') else ('w' if t.startswith('WARPSYNC') else ('b' if t.startswith(('BRA','BSYNC','BSSY')) else ('T' if 'TM' in t.split()[0] else ('S' if t.startswith('SYNCS') else '.')))) for a,t in best)
for i in range(0,len(line),100): print(line[i:i+100])
PY
