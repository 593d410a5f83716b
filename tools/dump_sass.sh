#!/bin/bash
# SASS listings of the two headline kernel variants into profiles/ (names carry a per-build hash: look them up).
# usage: tools/dump_sass.sh <tag>
set -eu
TAG=${1:-r01b}
LIB=raytracingdiffusioncurves_b200/librdc_b200.so
name() { cuobjdump -sass $LIB 2>/dev/null | grep "Function :" | grep "k_renderIL$1" | head -1 | awk '{print $3}'; }
cuobjdump -sass -fun "$(name b1ELb0ELb0ELi1E)" $LIB 2>/dev/null | grep -v "not found" > profiles/${TAG}_k_render_table_noportal.sass
cuobjdump -sass -fun "$(name b0ELb0ELb0ELi2E)" $LIB 2>/dev/null | grep -v "not found" > profiles/${TAG}_k_render_local_noportal.sass
cuobjdump -sass -fun "$(name b0ELb0ELb0ELi3E)" $LIB 2>/dev/null | grep -v "not found" > profiles/${TAG}_k_render_cut_noportal.sass
for f in profiles/${TAG}_k_render_table_noportal.sass profiles/${TAG}_k_render_local_noportal.sass profiles/${TAG}_k_render_cut_noportal.sass; do
  echo "$f: $(grep -cE '^\s+/\*[0-9a-f]{4}\*/' $f) instructions"
done
