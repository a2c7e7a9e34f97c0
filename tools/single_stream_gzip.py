"""Compress a file into ONE gzip member holding ONE deflate stream, in parallel, the way pigz does it: slices of the input
are deflated independently and end with a sync flush (an empty stored block, byte aligned), the last slice ends with the
stream's final block; one gzip header in front, one CRC-32 / ISIZE trailer behind.
    python tools/single_stream_gzip.py IN OUT [level] [slice_MB]"""
import os, struct, sys, zlib
from concurrent.futures import ProcessPoolExecutor

src, dst = sys.argv[1], sys.argv[2]
level = int(sys.argv[3]) if len(sys.argv) > 3 else 1
step = (int(sys.argv[4]) if len(sys.argv) > 4 else 64) << 20
size = os.path.getsize(src)


def piece(args):
    lo, hi, last = args
    with open(src, "rb") as f:
        f.seek(lo)
        data = f.read(hi - lo)
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    return co.compress(data) + (co.flush(zlib.Z_FINISH) if last else co.flush(zlib.Z_SYNC_FLUSH))


if __name__ == "__main__":
    jobs = [(lo, min(size, lo + step), lo + step >= size) for lo in range(0, max(size, 1), step)]
    with ProcessPoolExecutor(os.cpu_count()) as ex, open(dst, "wb") as out:
        out.write(b"\x1f\x8b\x08\x00\0\0\0\0\0\xff")
        for blob in ex.map(piece, jobs):
            out.write(blob)
        crc = 0
        with open(src, "rb") as f:
            while True:
                b = f.read(64 << 20)
                if not b:
                    break
                crc = zlib.crc32(b, crc)
        out.write(struct.pack("<II", crc & 0xffffffff, size & 0xffffffff))
