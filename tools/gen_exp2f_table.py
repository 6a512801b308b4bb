#!/usr/bin/env python3
"""Regenerate the 32-entry table of nc_expf (nanocall_b200/csrc/nc_device.cuh) from its definition:
T[i] = bits(RN_double(2^(i/32))) - (i << 47)."""
import struct
from decimal import Decimal, getcontext

getcontext().prec = 60
for i in range(32):
    d = float(Decimal(2) ** (Decimal(i) / Decimal(32)))
    bits = struct.unpack("<Q", struct.pack("<d", d))[0]
    print("0x%016xULL," % ((bits - (i << 47)) & 0xFFFFFFFFFFFFFFFF))
