"""The count-plane oracle (tests/counts_oracle.py) against the reference's own known answers
(T/MatchDatabase_Test.java:13-35 and :37-69), and the multi-rank reduction host logic under gloo (world size 2)."""
import os
import socket
import sys

import numpy as np
import pytest

import counts_oracle
import parity
from mapper_b200 import shard, synth

HERE = os.path.dirname(os.path.abspath(__file__))


def _codes(text):
    return np.array([synth.LETTERS.tobytes().index(c.encode()) for c in text], dtype=np.uint8)


def _results(alignments):
    """alignments: list (per query) of list (per component) of list (per choice) of list of (contig, reversed, blocks)."""
    r = dict(q_comp_off=[0], comp_choice_off=[0], choice_sa_off=[0], sa_block_off=[0], sa_contig=[], sa_reversed=[], blocks=[], q_status=[])
    for comps in alignments:
        r["q_status"].append(0)
        for choices in comps:
            for sas in choices:
                for contig, rev, blocks in sas:
                    r["sa_contig"].append(contig); r["sa_reversed"].append(rev)
                    for b in blocks:
                        r["blocks"].extend(b)
                    r["sa_block_off"].append(len(r["blocks"]) // 4)
                r["choice_sa_off"].append(len(r["sa_contig"]))
            r["comp_choice_off"].append(len(r["choice_sa_off"]) - 1)
        r["q_comp_off"].append(len(r["comp_choice_off"]) - 1)
    out = {k: np.asarray(v, dtype=np.int64) for k, v in r.items()}
    out["blocks"] = np.asarray(r["blocks"], dtype=np.int32)
    return out


def test_query_ending_with_mismatch():
    # T/MatchDatabase_Test.java:13-35 — depth 1 wherever the read agrees with the reference; the last base (T vs A) is an alternate
    ref = _codes("AACCACGA")
    q = _codes("AACCACGT")
    r = _results([[[[(0, 0, [[0, 0, 8, 8]])]]]])
    planes = counts_oracle.depth_planes([ref], [[q]], r, 0.0)
    total = planes[0].sum(axis=(0, 1))
    assert total.tolist() == [100] * 7 + [0]


def test_overlapping_paired_end_queries():
    # T/MatchDatabase_Test.java:37-69 — count == 1 at every reference position although the mates overlap
    ref = _codes("AACCACGATTAC")
    q1, q2 = _codes("AACCACGA"), _codes("CACGATTAC")
    r = _results([[[[(0, 0, [[0, 0, 8, 8]]), (0, 0, [[0, 3, 9, 9]])]]]])
    planes = counts_oracle.depth_planes([ref], [[q1, q2]], r, 0.0)
    assert planes[0].sum(axis=(0, 1)).tolist() == [100] * 12
    # weights: 100 where one mate covers, 50 + 50 in the overlap [3, 8)
    assert planes[0][0, 0].tolist() == [100] * 12


def test_near_query_end_region_and_choices():
    ref = _codes("ACGTACGTACGTACGTACGT")
    q = _codes("ACGTACGTAC")
    # two equally good choices (weight 1/2 each -> 50), end fraction 0.2 of a 10 bp read -> 2 bases at each end
    r = _results([[[[(0, 0, [[0, 0, 10, 10]])], [(0, 0, [[0, 4, 10, 10]])]]]])
    planes = counts_oracle.depth_planes([ref], [[q]], r, 0.2)
    assert planes[0][1, 0, :2].tolist() == [50, 50] and planes[0][0, 0, 2:4].tolist() == [50, 50]
    assert planes[0].sum() == 50 * 20


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _rank_main(rank, world, port, tmp):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, HERE)
    import xm_oracle as xo
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ref = synth.random_reference(60000, seed=81, n_contigs=2)
    db = xo.Oracle([(n, synth.codes_to_text(s)) for n, s in ref], sort_by_length=True, dup=dict(min_copies=2, window=1000))
    contigs = [db.contig(i) for i in range(db.num_contigs())]
    batch = synth.simulate_reads(contigs, 400, 100, seed=82, sub_rate=0.01, indel_rate=0.002, paired=True)
    lo, hi = shard.shard_bounds(len(batch["n_seqs"]), rank, world)
    mine = shard.take_shard(batch, lo, hi)
    res = db.align_batch(synth.DEFAULT_PARAMS, mine)
    reads = synth.unpack_reads(mine)
    planes = counts_oracle.depth_planes([c for _, c in contigs], reads, res, 0.1)
    flat = torch.from_numpy(np.concatenate([p.reshape(-1) for p in planes]).astype(np.int32))
    shard.allreduce_planes_(flat)
    if rank == 0:
        np.save(os.path.join(tmp, "reduced.npy"), flat.numpy())
        full = db.align_batch(synth.DEFAULT_PARAMS, batch)
        want = counts_oracle.depth_planes([c for _, c in contigs], synth.unpack_reads(batch), full, 0.1)
        np.save(os.path.join(tmp, "want.npy"), np.concatenate([p.reshape(-1) for p in want]).astype(np.int32))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_count_reduction(tmp_path):
    """world_size 2 over gloo: contiguous read shards (pairs kept together), per-rank planes, int32 all-reduce == single-process planes."""
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_rank_main, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got, want = np.load(tmp_path / "reduced.npy"), np.load(tmp_path / "want.npy")
    assert want.sum() > 0 and np.array_equal(got, want)


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 1000):
        for world in (1, 2, 3, 8):
            spans = [shard.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
