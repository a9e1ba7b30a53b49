#!/bin/bash
# Offline codegen check of lift_quad_kernel: registers / spills and the size of the pass-1 pixel loop (SASS
# instructions between the first DEPBAR.LE SB0,0x1 and the drain DEPBAR.LE SB0,0x0), which the kernel's issue-bound
# time follows closely.  Usage: tools/loop_stats.sh [path/to/liblm3d.so]
SO=${1:-/root/repo/3d-localisation-and-mapping_b200/lm3d/liblm3d.so}
cuobjdump -sass "$SO" | awk '/Function : /{f=$3} f ~ /lift_quad_kernel/ {print}' > /tmp/_quad.sass
a=$(grep -n "DEPBAR.LE SB0, 0x1" /tmp/_quad.sass | head -1 | cut -d: -f1)
b=$(grep -n "DEPBAR.LE SB0, 0x0" /tmp/_quad.sass | head -1 | cut -d: -f1)
n=$(sed -n "${a},${b}p" /tmp/_quad.sass | grep -cE "^\s+/\*[0-9a-f]{4}\*/")
bs=$(sed -n "${a},${b}p" /tmp/_quad.sass | grep -cE "BSSY|BSYNC")
echo "pass-1 loop: $n SASS instructions per group, $bs BSSY/BSYNC"
