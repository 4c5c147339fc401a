"""Two or more ranks on one node: PeerExchange (all-reduce + residual over NVLink peer memory) against the same arithmetic in
torch (fp32 accumulation in rank order, one rounding to fp16, residual added as a separate fp16 op), eagerly and replayed from a
CUDA graph, and the tensor-parallel decoder step against the NCCL exchange.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/check_peer_exchange.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mixq_b200.llama import CONFIGS, LlamaDecoder  # noqa: E402
from mixq_b200.tp import PeerExchange  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    M, H = 512, 4096
    two_shot = {"0": False, "1": True}.get(os.environ.get("MIXQ_TP_TWO_SHOT"), None)
    ex = PeerExchange(M, H, rank, world, two_shot=two_shot)
    kernel_copy = os.environ.get("CHECK_MEMCPY") != "1"       # a kernel (not a memcpy node) fills the partial buffer

    def reference(parts, res):                # == oracle.mixq_oracle.tp_exchange, restated in torch on the device
        acc = torch.zeros(M, H, dtype=torch.float32, device="cuda")
        for p_ in parts:                      # rank order, fp32 — what the kernel does (exact for 2 ranks, rounds beyond)
            acc += p_.float()
        return (acc.half().float() + res.float()).half()

    def compare(out, ref, what):
        bad = int((out != ref).sum())
        assert bad == 0, f"rank {rank} {what}: {bad} of {out.numel()} elements differ, max |diff| {float((out.float() - ref.float()).abs().max()):.3e}"

    def fill(dst, src_t):
        if kernel_copy:
            torch.add(src_t, 0, out=dst)
        else:
            dst.copy_(src_t)
    g = torch.Generator(device="cuda").manual_seed(1234)          # same stream of numbers on every rank
    for it in range(6):
        parts = [torch.randn(M, H, generator=g, device="cuda").half() for _ in range(world)]
        res = torch.randn(M, H, generator=g, device="cuda").half()
        fill(ex.next_partial(), parts[rank])
        out = ex.reduce(res).clone()
        compare(out, reference(parts, res), f"exchange {it}")
    # replayed from a graph (two exchanges per replay: the buffers alternate)
    src = [torch.zeros(M, H, dtype=torch.float16, device="cuda") for _ in range(2)]
    res = torch.randn(M, H, generator=g, device="cuda").half()
    outs = [None, None]
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for j in range(2):
            fill(ex.next_partial(), src[j])
            ex.reduce(res)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for j in range(2):
            fill(ex.next_partial(), src[j])
            outs[j] = ex.reduce(res).clone()
    for rep in range(5):
        allp = [[torch.randn(M, H, generator=g, device="cuda").half() for _ in range(world)] for _ in range(2)]
        for j in range(2):
            src[j].copy_(allp[j][rank])
        gr.replay()
        torch.cuda.synchronize()
        for j in range(2):
            compare(outs[j], reference(allp[j], res), f"graph replay {rep} exchange {j}")
    del gr
    # the decoder step: peer exchange vs NCCL all-reduce + add (both sum fp16 partials; orders differ -> tolerance)
    cfg = CONFIGS["tiny"] if world <= 4 else CONFIGS["tiny8"]
    tok = torch.randint(0, cfg.vocab, (256, 1), generator=torch.Generator().manual_seed(0)).cuda()
    logits = {}
    for mode in ("peer", "nccl"):
        os.environ["MIXQ_TP_EXCHANGE"] = mode
        m = LlamaDecoder(cfg, batch=256, bit=8, rank=rank, world_size=world)
        m.discover(tok)
        logits[mode] = m.step(tok).float()
        if m.xchg is not None:
            m.xchg.close()
    rel = float((logits["peer"] - logits["nccl"]).norm() / logits["nccl"].norm())
    assert rel < 5e-3, rel
    ex.close()
    dist.barrier()
    if rank == 0:
        print(f"peer exchange ok on {world} ranks ({'two' if ex.two_shot else 'one'}-shot, {'kernel' if kernel_copy else 'memcpy'} fill): bit-exact vs the torch restatement (eager + graph replay), decoder step rel diff vs NCCL {rel:.2e}")
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
