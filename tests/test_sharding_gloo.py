"""The N>1 path on CPU: world_size-2 gloo.  Each rank takes its newline-aligned shard of the same VCF
(bystro_vcf_b200.shard.partition), transforms it (here with the CPU oracle standing in for the GPU, which
this box lacks), and rank 0 reassembles the shard outputs in rank order.  No data-path collective: the only
communication is the final gather of outputs, exactly as in the multi-GPU host."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, hashlib
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import torch, torch.distributed as dist
from bystro_vcf_b200 import shard, synth, parse_preamble
from oracle import oracle as O
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank, world = dist.get_rank(), dist.get_world_size()
vcf = synth.header(11, 300) + synth.host_lines(11, 300, "chr1", 0, 900, 2)
w, chrom, off = parse_preamble(vcf)
lo, hi = shard.partition(vcf, off, len(vcf), world)[rank]
res = O.process_block(O.OracleConfig(), chrom, vcf[lo:hi])
outs = [None] * world
dist.all_gather_object(outs, (res.tsv, res.n_lines))
if rank == 0:
    whole = O.read_vcf(O.OracleConfig(), vcf)
    assert b"".join(o[0] for o in outs) == whole.tsv, "shard outputs in rank order != single-process output"
    assert sum(o[1] for o in outs) == whole.n_lines == 900
    assert all(o[1] > 0 for o in outs)
    print("OK", hashlib.md5(whole.tsv).hexdigest())
dist.barrier()
dist.destroy_process_group()
'''


def test_two_rank_sharding_reassembles_in_input_order(tmp_path):
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "OK" in outs[0]
