"""One rank of tests/test_gpu_shard.py::test_nccl_two_processes (also usable by hand on a multi-GPU box):

    python tests/shard_nccl_worker.py <rank> <nranks> <scratch dir>

Rank r takes GPU r, loads shard r of a small synthetic table, and runs its share of the proteins through
kg_batch_run_sharded over NCCL; the records must be byte-identical to the replicated-table run of the same sequences and
equal to the CPU oracle.  The communicator id travels through a file in the scratch directory."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main():
    rank, nranks, scratch = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    import kmergutsjava_b200 as kg
    from oracle import kgo
    from tests.parity import assert_same
    from tools import kg_synth as synth
    kgo.build()
    ctx = kg.Context(rank)
    idfile = os.path.join(scratch, "comm.id")
    if rank == 0:
        uid = kg.Comm.unique_id()
        with open(idfile + ".tmp", "wb") as f:
            f.write(uid)
        os.rename(idfile + ".tmp", idfile)
    else:
        t0 = time.time()
        while not os.path.exists(idfile):
            if time.time() - t0 > 120:
                raise SystemExit("no communicator id after 120 s")
            time.sleep(0.05)
        with open(idfile, "rb") as f:
            uid = f.read()
    comm = kg.Comm(ctx, rank, nranks, uid)
    u = synth.Universe(n_families=300, seed=0x4B470003)
    keys, otu, avg, fi, wt = u.signatures()
    img = synth.build_table_image(keys, otu, avg, fi, wt)
    shard = ctx.table_from_image_sharded(img, rank, nranks)
    full = ctx.table_from_image(img)
    otable = kgo.Table(data=img)
    seqs = u.proteins(6000, seed=9)
    mine = seqs[len(seqs) * rank // nranks: len(seqs) * (rank + 1) // nranks]
    # step 3 is three times the first batch: with the direct transport the exchange buffers (sized by the first step) overflow,
    # every rank sees it in its control block, and the buffers are re-made and re-mapped collectively before the step repeats
    for step, (mode, share) in enumerate([(kg.MODE_AA, mine), (kg.MODE_AA, mine[: 7 * rank]),
                                          (kg.MODE_DNA, [synth.genome(u, 20000, seed=40 + rank, index=i) for i in range(2)]),
                                          (kg.MODE_AA, mine * 3)]):
        sb, off = kgo.concat(share)
        params = kg.default_params(emit_hits=1)
        b = ctx.upload(mode, sb, off)
        res = comm.run(shard, b, params)
        rep = ctx.run(full, mode, sb, off, params)
        for name in ("hits", "calls", "otus"):
            assert getattr(res, name).tobytes() == getattr(rep, name).tobytes(), f"rank {rank} step {step}: {name}"
        ref = kgo.run(otable, kgo.make_params(aa=mode == kg.MODE_AA), sb, off, kgo.DIRECT_PROBE)
        assert_same(res, ref, what=f"rank {rank} step {step}")
        st = comm.stats
        assert st.keys_sent == ref.num_kmers and st.replies_received == len(ref.hits)
        print(f"rank {rank} step {step}: {len(share)} sequences, {st.keys_sent} keys ({st.keys_remote} remote), "
              f"{st.replies_received} hits, {len(ref.calls)} calls, {st.ms_total:.2f} ms", flush=True)
        res.free()
        rep.free()
        b.free()
    comm.free()
    shard.free()
    full.free()
    ctx.close()
    print("OK", flush=True)


if __name__ == "__main__":
    main()
