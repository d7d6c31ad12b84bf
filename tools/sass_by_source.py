#!/usr/bin/env python3
"""SASS instruction count of one kernel by source region: tools/sass_by_source.py <nvdisasm -g output> <first line> <last line>
(lines of the kernel's .text section in that file)."""
import re, sys, collections, bisect
lines = open(sys.argv[1]).read().split('\n')[int(sys.argv[2]):int(sys.argv[3])]
cur = None; cnt = collections.Counter()
for l in lines:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+\S', l) and cur: cnt[cur] += 1
byfile = collections.Counter()
for (f, ln), c in cnt.items(): byfile[f] += c
print(sum(cnt.values()), byfile.most_common())
starts = [(42,'dot3'),(48,'to_node'),(81,'rcp'),(92,'make_invdir'),(101,'div_hoisted'),(116,'slab_fast'),(141,'numer'),(153,'slab'),(207,'mesh_invdir'),(218,'sphere_hit'),(244,'plane_hit'),(263,'tri_hit'),(293,'load_pair'),(305,'bvh_walk'),(397,'occ_setup'),(424,'occ_node'),(455,'ref_reaches'),(482,'occ_candidate'),(507,'occ_walk'),(546,'occ_candidate_closest'),(567,'occ_walk_closest'),(597,'lbvh_walk'),(646,'mesh_hit'),(662,'sphere_or_plane'),(688,'object_hit'),(697,'load_node'),(715,'bound_culled'),(731,'local_ray_of'),(751,'top'),(908,'finalize')]
fc = collections.Counter()
for (f, ln), c in cnt.items():
    if f == 'intersect.cuh':
        fc[starts[bisect.bisect_right([s for s, _ in starts], ln) - 1][1]] += c
print(fc.most_common())
for fn in ('rtu_kernels.cu', 'camera.cuh', 'shade.cuh'):
    kc = collections.Counter()
    for (f, ln), c in cnt.items():
        if f == fn: kc[ln // 20 * 20] += c
    print(fn, sorted(kc.items()))
