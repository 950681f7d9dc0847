#!/bin/bash
# sass_count.sh LIB KERNEL_SUBSTRING [top] -> opcode histogram of one kernel's SASS (static counts)
cuobjdump -sass "$1" | awk -v k="$2" '/Function :/ {on = index($0, k) > 0} on && /\/\*[0-9a-f][0-9a-f][0-9a-f][0-9a-f]\*\/ / {print}' > /tmp/_k.sass
echo "total $(wc -l < /tmp/_k.sass)"
sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+(@!?U?P[0-9T]+ )?//' /tmp/_k.sass | awk '{split($1,a,"."); c[a[1]]++} END {for (o in c) print c[o], o}' | sort -rn | head -${3:-25} | tr '\n' ';'
echo
