#!/usr/bin/env python3
"""Dump float/double constants from this box's glibc libm.so.6 by virtual address.

Test infrastructure only: used while restating glibc 2.39's flt-32 routines
(the third-party libm the reference reaches through std::sin/cos/atan2/...).
"""
import struct, subprocess, sys
LIBM = '/lib/x86_64-linux-gnu/libm.so.6'
data = open(LIBM, 'rb').read()
out = subprocess.run(['readelf', '-lW', LIBM], capture_output=True, text=True).stdout
segs = []
for l in out.splitlines():
    p = l.split()
    if p and p[0] == 'LOAD':
        segs.append((int(p[1], 16), int(p[2], 16), int(p[4], 16)))
def rd(va, n):
    for off, v, sz in segs:
        if v <= va < v + sz:
            return data[off + va - v: off + va - v + n]
    raise KeyError(hex(va))
def f32(va):
    u = struct.unpack('<I', rd(va, 4))[0]
    return u, struct.unpack('<f', rd(va, 4))[0]
if __name__ == '__main__':
    for a in sys.argv[1:]:
        if a.startswith('d:'):
            va = int(a[2:], 16); print(hex(va), struct.unpack('<d', rd(va, 8))[0].hex())
        else:
            va = int(a, 16); u, f = f32(va); print(hex(va), hex(u), float(f).hex(), repr(f))
