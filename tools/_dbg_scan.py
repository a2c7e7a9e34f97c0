import random, sys
import numpy as np
sys.path.insert(0, ".")
import oracle
from vfind_b200 import api
sys.path.insert(0, "tests")
import test_gpu_parity as T
pre, suf = b"GGGCCCAGCCGGCCGGATTA", b"CCGGAGGCGGAGGTTCAGAC"
rng = random.Random(len(pre) * 7 + 3)
rnd = lambda n: bytes(rng.choice(b"ACGT") for _ in range(n))
seqs = []
for i in range(6000):
    k = rng.randrange(12)
    if k == 0:
        seqs.append(rnd(rng.randrange(0, 9)) + pre[:rng.randrange(8, len(pre))] + b"N" + rnd(rng.randrange(0, 20)) + pre +
                    rnd(30) + suf[:rng.randrange(8, len(suf))] + rnd(3) + suf + rnd(rng.randrange(0, 9)))
    elif k == 1:
        seqs.append(rnd(rng.randrange(0, 17)) + pre + rnd(rng.randrange(0, 40)) + pre + rnd(21) + suf + rnd(rng.randrange(0, 40)) + suf)
    elif k == 2:
        cut = rng.randrange(1, len(pre))
        seqs.append(rnd(rng.randrange(0, 60)) + pre[:cut])
        seqs.append(pre[cut:] + rnd(rng.randrange(0, 60)) + suf[:rng.randrange(1, len(suf))])
    elif k == 3:
        seqs.append(b"" if rng.random() < 0.5 else rnd(rng.randrange(1, len(pre))))
    elif k == 4 and i % 50 == 0:
        seqs.append(rnd(rng.randrange(3000, 9000)) + pre + rnd(300) + suf + rnd(rng.randrange(0, 3000)))
    elif k == 5:
        seqs.append(suf + pre[:5] + pre + rnd(12))
    else:
        seqs.append(rnd(rng.randrange(0, 13)) + pre + rnd(rng.choice((21, 24, 30, 198))) + suf + rnd(rng.randrange(0, 13)))
text, off, ln = oracle.pack_reads(seqs)
want, odiag, _ = oracle.process_reads(oracle.make_params((pre, suf), accept_prefix_alignment=1.0, accept_suffix_alignment=1.0,
                                                         skip_translation=True), text, off, ln, want_diag=True)
n = len(off)
orders = {"random": np.array(rng.sample(range(n), n)), "seventh": (lambda a: ([a.__setitem__(i, n-1-i) or a.__setitem__(n-1-i, i) for i in range(0, n//2, 7)], a)[1])(np.arange(n)),
          "reverse": np.arange(n)[::-1].copy()}
for name, order in orders.items():
    got2, diag2 = T._scan_only(text, off[order], ln[order], (pre, suf))
    for f in ("exact_prefix", "exact_suffix", "start", "end"):
        bad = np.nonzero(diag2[f] != odiag[f][order])[0]
        print(name, f, "bad", len(bad), bad[:10])
        for b in bad[:4]:
            r = order[b]
            print("   pos", b, "unit", b // 32, "lane", b % 32, "read", r, "off", off[r], "len", ln[r], "got", diag2[f][b], "want", odiag[f][r], seqs[r][:80])
    print(name, "table equal", got2 == want)
