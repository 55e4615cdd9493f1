#!/usr/bin/env python
"""Amplicon windows (1-based closed, primers included) from a primer scheme BED of the reference tree:
usage: python wepp_b200/data/make_amplicons.py /root/reference/primers/ARTICv4_1.bed > wepp_b200/data/ARTICv4_1.amplicons.tsv
One line per amplicon: index, start, end.  An amplicon runs from the leftmost base of its LEFT primers to the
rightmost base of its RIGHT primers (alternate primers included).  The BED files are the reference's public primer
schemes (primers/*.bed); only the derived coordinates are kept here."""
import re, sys
amp = {}
for line in open(sys.argv[1]):
    f = line.split()
    if len(f) < 4:
        continue
    m = re.search(r"_(\d+)_(LEFT|RIGHT)", f[3])
    if not m:
        continue
    k, side = int(m.group(1)), m.group(2)
    lo, hi = min(int(f[1]), int(f[2])) + 1, max(int(f[1]), int(f[2]))   # BED: 0-based half-open (some schemes list RIGHT primers reversed)
    a = amp.setdefault(k, [10 ** 9, 0])
    if side == "LEFT":
        a[0] = min(a[0], lo)
    else:
        a[1] = max(a[1], hi)
print("# amplicon\tstart\tend   (1-based closed; derived from " + sys.argv[1].split("/")[-1] + ")")
for k in sorted(amp):
    if amp[k][0] < amp[k][1]:
        print(f"{k}\t{amp[k][0]}\t{amp[k][1]}")
