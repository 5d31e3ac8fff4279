"""TEST INFRASTRUCTURE (runs the compiled reference under oracle/_ref next to the product).
Timing of the 8(f) rows on the GPU box: analogy evaluator at the Google-set shape and the host
tokenizer.  python tests/tools/eval_perf.py [V] [D] [questions]"""
import os, sys, time, subprocess, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import word2bits_b200 as w2b
import bench

V = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
D = int(sys.argv[2]) if len(sys.argv) > 2 else 800
NQ = int(sys.argv[3]) if len(sys.argv) > 3 else 19544
tmp = tempfile.mkdtemp()
rng = np.random.default_rng(0)
vf = os.path.join(tmp, "vec.bin")
t0 = time.time()
with open(vf, "wb") as f:
    f.write(b"%d %d\n" % (V, D))
    block = 20000
    for a in range(0, V, block):
        n = min(block, V - a)
        x = np.where(rng.random((n, D)) < 0.5, -1.0 / 3, 1.0 / 3).astype(np.float32)  # 1-bit vectors
        for i in range(n):
            f.write(b"w%d " % (a + i) + x[i].tobytes() + b"\n")
print("vector file %.1f GB written in %.1f s" % (os.path.getsize(vf) / 1e9, time.time() - t0), flush=True)
qf = os.path.join(tmp, "q.txt")
with open(qf, "w") as f:
    for s in range(14):
        f.write(": s%d\n" % s)
        for _ in range(NQ // 14):
            f.write(" ".join("w%d" % i for i in rng.integers(0, min(V, 30000), 4)) + "\n")
t0 = time.time()
rep, acc = w2b.compute_accuracy(vf, qf, bitlevel=1)
wall = time.time() - t0
flops = 2.0 * acc["questions_seen"] * acc["vocab"] * acc["size"]
print("GPU evaluator: %d questions x V=%d x D=%d: kernels %.1f ms (%.1f TFLOP/s of contraction), wall %.1f s (file read + H2D included); "
      "tensor-core filter let %.1f candidates per question through, %.1f re-scored in fp32"
      % (acc["questions_seen"], acc["vocab"], acc["size"], acc["gpu_ms"], flops / acc["gpu_ms"] / 1e9, wall,
         acc["candidates"] / max(acc["questions_seen"], 1), acc["rescored"] / max(acc["questions_seen"], 1)), flush=True)
os.environ["W2B_EVAL_SIMT"] = "1"
rep2, acc2 = w2b.compute_accuracy(vf, qf, bitlevel=1)
del os.environ["W2B_EVAL_SIMT"]
print("same with every score in fp32 on the SIMT cores (W2B_EVAL_SIMT=1): kernels %.1f ms; reports identical: %s"
      % (acc2["gpu_ms"], rep == rep2), flush=True)
if len(sys.argv) > 4 and sys.argv[4] == "eval-only":
    sys.exit(0)
refbin = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "oracle", "_ref", "compute_accuracy")
if os.path.exists(refbin):
    qs = os.path.join(tmp, "qs.txt")
    nsmall = 40
    with open(qs, "w") as f:
        f.write(": s0\n")
        for line in open(qf).read().splitlines()[1:1 + nsmall]:
            f.write(line + "\n")
    t0 = time.time()
    subprocess.run([refbin, vf, "1", "0"], stdin=open(qs), capture_output=True)
    t_all = time.time() - t0
    t0 = time.time()
    subprocess.run([refbin, vf, "1", "0"], stdin=open(os.devnull), capture_output=True)
    t_load = time.time() - t0
    per_q = (t_all - t_load) / nsmall
    print("reference compute_accuracy (1 thread): load %.1f s, %.3f s per question -> %.0f s for %d questions"
          % (t_load, per_q, per_q * acc["questions_seen"], acc["questions_seen"]), flush=True)
# ---- host tokenizer
cdf, _ = bench.zipf_cdf(400000)
ids = bench.synth_ids(30_000_000, 7, cdf)
path = bench._write_text(ids, "w2b_tok_")
for th in ("1", str(os.cpu_count())):
    os.environ["W2B_TOKENIZER_THREADS"] = th
    t0 = time.time()
    c = w2b.Corpus(path, 5)
    dt = time.time() - t0
    print("tokenizer %s threads: %.2f s, %.1f M words/s (V=%d)" % (th, dt, c.train_words / dt / 1e6, c.vocab_size), flush=True)
    c.close()
os.unlink(path)
