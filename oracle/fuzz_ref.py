"""Fuzz campaign (test infrastructure): the oracle's restated amd64 flavour against the reference's
real assembly (oracle/_ref), all three levels, plus both decoders on every stream and on mutated
streams.  Not part of the test suite (minutes of CPU); run by hand in the build container:

    python oracle/fuzz_ref.py 320        # 320 seeds x 150 inputs x 3 levels

Last run (round 1): 144 000 encoder comparisons, ~290 000 mutated-stream decodes, 0 differences.
"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from multiprocessing import Pool

def gen(rng):
    import patterns
    return patterns.random_structure(rng)


def work(seed):
    from oracle import binding as o, refasm
    rng = np.random.default_rng(seed)
    bad = []
    for it in range(150):
        d = gen(rng)
        for lv in (-1, 1, 2):
            a = refasm.encode_block(d, lv); b = o.encode_block(d, lv, flavor="asm")
            if a != b:
                bad.append((seed, it, lv, d.size, len(a), len(b)))
                np.save("/tmp/fuzz_bad_%d_%d_%d.npy" % (seed, it, lv), d)
            elif a:
                st, out = o.decode_block(a, d.size)
                st2, out2 = refasm.decode_block(a, d.size)
                if st or st2 or out != d.tobytes() or out2 != out:
                    bad.append((seed, it, lv, "decode", st, st2))
                if d.size <= 100000:
                    m0 = np.frombuffer(a, dtype=np.uint8)
                    for k in range(3):
                        m = m0.copy()
                        if k == 2 and m.size > 2:
                            m = m[:int(rng.integers(1, m.size))]
                        else:
                            for q in rng.integers(0, m.size, 1 + k):
                                m[q] = rng.integers(0, 256)
                        s1, o1 = o.decode_block(m, d.size); s2, o2 = refasm.decode_block(m, d.size)
                        if (s1 != 0) != (s2 != 0) or (s1 == 0 and o1 != o2):
                            bad.append((seed, it, lv, "mutdecode", s1, s2))
                            np.save("/tmp/fuzz_badmut_%d_%d_%d_%d.npy" % (seed, it, lv, k), m)
    return bad

if __name__ == "__main__":
    t0 = time.time()
    with Pool(8) as p:
        res = p.map(work, range(5000, 5000 + int(sys.argv[1])))
    bad = [b for r in res for b in r]
    print("cases", int(sys.argv[1]) * 150 * 3, "bad", len(bad), bad[:10], "%.0f s" % (time.time() - t0))
