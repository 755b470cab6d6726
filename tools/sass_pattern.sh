#!/bin/bash
# print the LDS/DADD issue pattern of the J=16 fold loops in the mbest kernel (L = load, A = add)
cuobjdump -sass -fun '_ZN2pp12mbest_kernelEPKdliiiiiiiiiNS_6TablesEPjPdS4_PiS5_S4_S4_Py' ${1:-pyperiod_b200/libpyperiod_b200.so} 2>/dev/null > /tmp/mb.sass
python - <<'PY'
import re
ins=[l for l in open('/tmp/mb.sass').read().split('\n') if re.match(r'\s+/\*[0-9a-f]{4,5}\*/',l)]
ops=[re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);',l).group(2) for l in ins]
idx=[i for i,o in enumerate(ops) if 'LDS.64' in o and '+0xf00]' in o]
seen=0
for i in idx:
    # find enclosing backward branch after i
    seg=ops[i-20:i+80]
    pat=''.join('L' if 'LDS.64' in o else 'A' if o.startswith('DADD') else 'B' if 'BRA' in o else '.' for o in seg)
    print(pat)
    seen+=1
    if seen>=6: break
print('total instrs', len(ops))
PY
