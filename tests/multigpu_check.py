"""Run under torchrun with 2+ ranks (tests/test_gpu_multi.py does that):
every rank trains its shard of streams for a few steps with the deltas summed
across ranks, once through NCCL and once through the fused peer-memory
kernel; rank 0 replays ALL streams of ALL ranks on the CPU oracle and prints
the worst relative weight difference of each variant as one JSON line."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import oracle
    from recur_b200 import api, abi, dist as rdist
    from helpers import make_net, weights, arr, fptr, u8ptr, markov_text, rel_err

    rank, world, local = rdist.env_rank()
    torch.cuda.set_device(local)
    L = api.load_library()
    L.rnn_b200_set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rdist.join_comm(L, dist, rank, world, device="cuda")

    n, steps, lr = 64, 5, 1e-4
    shape = dict(input_size=42, hidden=127, output=42, depth=12)
    text = markov_text(4000, 42, seed=2)
    lo, hi = rdist.shard_bounds(len(text), rank, world)
    my_text = np.ascontiguousarray(text[lo:hi])
    results = {}
    for variant in ("nccl", "p2p"):
        net = make_net(L, seed=1, lr=lr, **shape)
        ih0, ho0 = [w.copy() for w in weights(net)]
        nets = L.rnn_new_training_set(net, n)
        batch = L.rnn_batch_new(nets, n)
        if variant == "p2p":
            hb = (C.c_uint8 * 192)()
            assert L.rnn_batch_p2p_export(batch, hb) == 0
            mine = torch.tensor(list(hb), dtype=torch.uint8, device="cuda")
            allh = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(allh, mine)
            raw = b"".join(bytes(t.cpu().tolist()) for t in allh)
            buf = (C.c_uint8 * len(raw)).from_buffer_copy(raw)
            assert L.rnn_batch_p2p_attach(batch, buf, rank, world) == 0
        L.rnn_batch_text_upload(batch, u8ptr(my_text), len(my_text))
        L.rnn_batch_text_train(batch, 0, steps, 0, 0.9, 0.0, None)
        L.rnn_b200_synchronize()
        ih, ho = [w.copy() for w in weights(net)]
        # every replica must hold the same weights
        t = torch.tensor(ih, device="cuda")
        tmax, tmin = t.clone(), t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        results[variant] = dict(ih=ih, ho=ho, replicas_equal=bool(torch.equal(tmax, tmin)),
                                ih0=ih0, ho0=ho0)
        L.rnn_batch_delete(batch)
        L.rnn_delete_training_set(nets, n, 0)
        dist.barrier()
    # ---- a bottom layer in front (the gstclassify / parrot shape): its deltas
    # are summed over the ranks too
    nb, F, bsteps = 8, 12, 5
    bshape = (F, 7, 67, 4)   # n_inputs, r_inputs, hidden, output
    from helpers import STD_FLAGS

    def make_bottom(lib):
        net = lib.rnn_new_with_bottom_layer(bshape[0], bshape[1], bshape[2], bshape[3], STD_FLAGS,
                                            8, None, 4, 0.01, 0.9, 0.0, abi.RNN_RELU, 0)
        lib.rnn_randomise_weights_auto(net)
        return net

    def all_weights(net):
        bl = net.contents.bottom_layer.contents
        return [w.copy() for w in weights(net)] + [arr(bl.weights, bl.i_size * bl.o_size).copy()]

    rs = np.random.RandomState(5)
    feats = rs.random_sample((bsteps, world * nb, F)).astype(np.float32)
    tgt = rs.randint(0, 4, size=(bsteps, world * nb)).astype(np.uint8)
    bnet = make_bottom(L)
    bnets = L.rnn_new_training_set(bnet, nb)
    bbatch = L.rnn_batch_new(bnets, nb)
    mine = slice(rank * nb, (rank + 1) * nb)
    for t in range(bsteps):
        L.rnn_bptt_clear_deltas(bnet) if t % 2 == 0 else None   # the accumulator grows on odd steps
        L.rnn_batch_set_inputs(bbatch, fptr(np.ascontiguousarray(feats[t, mine])))
        L.rnn_batch_opinion(bbatch, 0.0)
        L.rnn_batch_softmax_error(bbatch, u8ptr(np.ascontiguousarray(tgt[t, mine])), None, None)
        L.rnn_batch_calc_deltas(bbatch, 0)
        L.rnn_batch_advance(bbatch)
        L.rnn_apply_learning(bnet, abi.RNN_MOMENTUM_NESTEROV, 0.9)
    L.rnn_b200_synchronize()
    bgot = all_weights(bnet)
    tb = torch.tensor(bgot[2], device="cuda")
    bmax, bmin = tb.clone(), tb.clone()
    dist.all_reduce(bmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(bmin, op=dist.ReduceOp.MIN)
    bottom_equal = bool(torch.equal(bmax, bmin))
    L.rnn_batch_delete(bbatch)
    dist.barrier()
    # ---- the cell automaton with its rows shared between the ranks: every
    # rank's gathered picture equals the one a single GPU computes
    W, Hh = 24, 8 * world
    off_y = np.array([(dx, dy) for dy in range(-2, 3) for dx in range(-2, 3)
                      if abs(dx) + abs(dy) <= 2 or (abs(dx), abs(dy)) == (2, 2)][:17], dtype=np.int32)
    off_c = np.array([(dx, dy) for dy in (-1, 0, 1) for dx in (-1, 0, 1) if (dx, dy) != (0, 0)],
                     dtype=np.int32)
    cnet = make_net(L, input_size=35, hidden=51, output=3, depth=10, seed=11, lr=3e-3)
    whole = L.rnn_cells_new(cnet, W, Hh)
    banded = L.rnn_cells_new_sharded(cnet, W, Hh)
    u8p, ip = C.POINTER(C.c_uint8), C.POINTER(C.c_int)
    frame = np.random.RandomState(9).randint(0, 256, size=3 * W * Hh).astype(np.uint8)
    start = frame.copy()
    cells_ok = bool(whole) and bool(banded)
    for f in range(3):
        a_out, b_out = np.zeros_like(frame), np.zeros_like(frame)
        for obj, out_ in ((whole, a_out), (banded, b_out)):
            L.rnn_cells_rnnca_frame(obj, frame.ctypes.data_as(u8p), out_.ctypes.data_as(u8p),
                                    off_y.ctypes.data_as(ip), len(off_y), off_c.ctypes.data_as(ip),
                                    len(off_c), 2, f % 2)
        cells_ok = cells_ok and bool(np.array_equal(a_out, b_out)) and bool(a_out.any())
        frame = a_out
    L.rnn_cells_forget(banded)
    ran = np.zeros_like(frame)
    for f in range(3):   # the same three frames, the pictures staying on the device
        L.rnn_cells_rnnca_run(banded, start.ctypes.data_as(u8p) if f == 0 else None, 1,
                              ran.ctypes.data_as(u8p), off_y.ctypes.data_as(ip), len(off_y),
                              off_c.ctypes.data_as(ip), len(off_c), 2, f % 2)
    cells_ok = cells_ok and bool(np.array_equal(ran, frame))
    # several frames in one call: between them the ranks swap only the rows
    # next to their bands; wrapped and clamped edges; then on from there
    for edges in (0, 1):
        pics = []
        for obj in (whole, banded):
            L.rnn_cells_forget(obj)
            p1, p2 = np.zeros_like(frame), np.zeros_like(frame)
            L.rnn_cells_rnnca_run(obj, start.ctypes.data_as(u8p), 4, p1.ctypes.data_as(u8p),
                                  off_y.ctypes.data_as(ip), len(off_y), off_c.ctypes.data_as(ip),
                                  len(off_c), 2, edges)
            L.rnn_cells_rnnca_run(obj, None, 3, None, off_y.ctypes.data_as(ip), len(off_y),
                                  off_c.ctypes.data_as(ip), len(off_c), 2, edges)
            L.rnn_cells_rnnca_run(obj, None, 2, p2.ctypes.data_as(u8p), off_y.ctypes.data_as(ip),
                                  len(off_y), off_c.ctypes.data_as(ip), len(off_c), 2, edges)
            pics.append((p1, p2))
        cells_ok = cells_ok and bool(np.array_equal(pics[0][0], pics[1][0])) \
            and bool(np.array_equal(pics[0][1], pics[1][1])) and bool(pics[0][1].any()) \
            and not bool(np.array_equal(pics[0][0], pics[0][1]))
    tc = torch.tensor([1 if cells_ok else 0], device="cuda")
    dist.all_reduce(tc, op=dist.ReduceOp.MIN)
    cells_ok = bool(tc.item())
    L.rnn_cells_delete(whole)
    L.rnn_cells_delete(banded)
    dist.barrier()
    bottom = None
    if rank == 0 and oracle.have_ref():
        ref = oracle.load_ref(strict=True)
        r = make_bottom(ref)
        rn = ref.rnn_new_training_set(r, world * nb)
        for t in range(bsteps):
            if t % 2 == 0:
                ref.rnn_bptt_clear_deltas(r)
            for j in range(world * nb):
                c = rn[j].contents
                out = ref.rnn_opinion(rn[j], fptr(feats[t, j]), 0.0)
                ref.ref_softmax_best_guess(c.bptt.contents.o_error, out, 4)
                arr(c.bptt.contents.o_error, c.o_size)[tgt[t, j]] += 1.0
                # rnn_bptt_calc_deltas(net, j ? 1 : 0): the deltas start afresh, the
                # bottom layer's shared error accumulator does not (recur-nn.c:751-755)
                ref.rnn_bptt_calc_deltas(rn[j], 1 if j else 0, None)
                ref.rnn_bptt_advance(rn[j])
            ref.rnn_apply_learning(r, abi.RNN_MOMENTUM_NESTEROV, 0.9)
        want = all_weights(r)
        bottom = {"rel": [rel_err(g, w) for g, w in zip(bgot, want)], "replicas_equal": bottom_equal}
    if rank == 0:
        port = oracle.load_port()
        s = port.oracle_set_new(42, 127, 42, n * world, 12, lr, abi.RNN_RELU, 1,
                                fptr(results["nccl"]["ih0"]), fptr(results["nccl"]["ho0"]))
        for i in range(steps):
            cur, nxt = [], []
            for r in range(world):
                a, b = rdist.shard_bounds(len(text), r, world)
                t_r = text[a:b]
                for p in rdist.stream_positions(len(t_r), n, i):
                    cur.append(t_r[p])
                    nxt.append(t_r[p + 1])
            cur = np.array(cur, dtype=np.uint8)
            nxt = np.array(nxt, dtype=np.uint8)
            port.oracle_set_char_step(s, u8ptr(cur), u8ptr(nxt), 0.9, None, None, None)
        want_ih = arr(port.oracle_set_wih(s), len(results["nccl"]["ih"])).copy()
        want_ho = arr(port.oracle_set_who(s), len(results["nccl"]["ho"])).copy()
        out = {"world": world}
        for variant, r in results.items():
            out[variant] = {"ih_rel": rel_err(r["ih"], want_ih), "ho_rel": rel_err(r["ho"], want_ho),
                            "replicas_equal": r["replicas_equal"]}
        out["bottom"] = bottom
        out["cells_sharded_equal"] = cells_ok
        print("MULTIGPU_CHECK " + json.dumps(out))
    dist.barrier()
    L.rnn_b200_comm_leave()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
