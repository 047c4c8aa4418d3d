import sys
sys.path.insert(0, ".")
from zero_chain_b200 import groth16 as zk
ctx = zk.Context(0)
names = {10: "mul interleaved", 11: "mul separated (Comba+REDC)", 12: "sqr via interleaved mul", 13: "sqr separated", 14: "a*b - c*d: two products", 15: "a*b - c*d: shared reduction"}
for threads in (128, 256):
    for f in (10, 11, 12, 13, 14, 15):
        per_s, ms = zk.bench_modmul(ctx, f, 148 * 4, threads, 2000)
        print("threads=%d %-34s %.3e ops/s (%.2f ms)" % (threads, names[f], per_s, ms), flush=True)
