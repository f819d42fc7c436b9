#!/bin/bash
# static SASS instruction count per kernel of the in-tree library (a quick proxy for code bloat before spending GPU time)
cuobjdump -sass "${1:-vqacl_b200/libvqacl_b200.so}" 2>/dev/null | awk '/Function :/ {name=$3} /^ +\/\*[0-9a-f]+\*\/ +[A-Z@]/ {cnt[name]++} END {for (n in cnt) print cnt[n], n}' | sort -rn
