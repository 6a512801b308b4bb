"""Writers for the event-table inputs of the nanocall-b200 CLI (see host/pipeline.cpp for the readers).

.events.tsv  one read: optional "#read_id <id>" line, then rows "strand mean stdv start length"
             (the last four columns are what the reference's Event::operator>> reads, Event.hpp:59-68)
.ncev        many reads: "NCEV0001", u32 n_reads, per read u32 id_len, id, u32 n0, u32 n1, then for each
             strand mean[n] stdv[n] start[n] length[n] as little-endian float32
.fast5 (NCRW0001)  raw event tables: "NCRW0001", u32 n_reads, per read u32 id_len, id, f64 sampling_rate,
             u32 n_events, then n_events x {f64 mean, f64 stdv, i64 start, i64 length} -- the EventDetection events
             of a fast5 file as fast5::File::get_eventdetection_events returns them (fast5.hpp:55-68); one read per
             file stands in for one fast5 file
"""
import struct

import numpy as np


def write_events_tsv(path, read_id, strands):
    """strands: list (index = strand) of dict(mean, stdv, start[, length]) or None."""
    with open(path, "w") as f:
        f.write(f"#read_id\t{read_id}\n")
        for st, ev in enumerate(strands):
            if ev is None:
                continue
            length = ev.get("length", np.zeros_like(ev["mean"]))
            for m, s, t, l in zip(ev["mean"], ev["stdv"], ev["start"], length):
                f.write(f"{st}\t{float(m):.9g}\t{float(s):.9g}\t{float(t):.9g}\t{float(l):.9g}\n")


def write_ncev(path, reads):
    """reads: list of (read_id, [strand0 dict or None, strand1 dict or None])."""
    with open(path, "wb") as f:
        f.write(b"NCEV0001")
        f.write(struct.pack("<I", len(reads)))
        for read_id, strands in reads:
            rid = read_id.encode()
            n = [0 if (len(strands) <= st or strands[st] is None) else len(strands[st]["mean"]) for st in range(2)]
            f.write(struct.pack("<I", len(rid)))
            f.write(rid)
            f.write(struct.pack("<II", n[0], n[1]))
            for st in range(2):
                if n[st] == 0:
                    continue
                ev = strands[st]
                length = ev.get("length", np.zeros(n[st], np.float32))
                for a in (ev["mean"], ev["stdv"], ev["start"], length):
                    f.write(np.ascontiguousarray(a, dtype="<f4").tobytes())


RAW_DTYPE = np.dtype([("mean", "<f8"), ("stdv", "<f8"), ("start", "<i8"), ("length", "<i8")])


def write_ncrw(path, reads):
    """reads: list of (read_id, sampling_rate, structured array with RAW_DTYPE)."""
    with open(path, "wb") as f:
        f.write(b"NCRW0001")
        f.write(struct.pack("<I", len(reads)))
        for read_id, rate, ev in reads:
            rid = read_id.encode()
            f.write(struct.pack("<I", len(rid)))
            f.write(rid)
            f.write(struct.pack("<d", float(rate)))
            f.write(struct.pack("<I", len(ev)))
            f.write(np.ascontiguousarray(ev, dtype=RAW_DTYPE).tobytes())
