"""CPU restatement of the row-sharded exchange's device ops (include/tt_b200.h section 7 and
tt_emb_segment_grad_lists), written from the wire-format description in the header, NOT from the kernels.

TEST INFRASTRUCTURE ONLY: injected as ``dev_ops`` into sharded.ShardedTableGroup by the gloo world-size-2 tests (host
logic on CPU) and used as the checker of the CUDA kernels in the ``-m gpu`` tests.  The product never imports it.
Plain loops over samples / entries: meant for small cases.
"""
import torch

POOL_NONE, POOL_SUM, POOL_MEAN = 0, 1, 2


class CpuShardOps:
    @staticmethod
    def route(ids, pad, vocab, world, send, block_ints, off_base, rows_base, cap, n_pad, flags):
        B, L = ids.shape
        send[:, rows_base:rows_base + cap] = -1
        counts = torch.zeros(world, B, dtype=torch.int64)
        lists = [[] for _ in range(world)]
        for b in range(B):
            pads = 0
            for l in range(L):
                i = int(ids[b, l])
                if pad is not None and i == pad:
                    pads += 1
                elif i < 0 or i >= vocab:
                    flags[0] |= 1
                else:
                    w = i % world
                    counts[w, b] += 1
                    lists[w].append(i // world)
            if n_pad is not None:
                n_pad[b] = pads
        for w in range(world):
            off = torch.zeros(B + 1, dtype=torch.int64)
            off[1:] = torch.cumsum(counts[w], 0)
            send[w, off_base:off_base + B + 1] = off.to(torch.int32)
            if int(off[B]) > cap:
                flags[0] |= 2
            rows = torch.tensor(lists[w][:cap], dtype=torch.int32)
            send[w, rows_base:rows_base + rows.numel()] = rows

    @staticmethod
    def owner_gather(table, local_rows, world, recv, block_ints, off_base, rows_base, cap, n_rows, pooled, out, block_floats,
                     vec_base, pos_src):
        D = table.shape[1]
        tab = table.float()
        for s in range(world):
            if pooled:
                for b in range(n_rows):
                    e0 = int(recv[s, off_base + b])
                    e1 = min(int(recv[s, off_base + b + 1]), cap)
                    e0 = min(e0, e1)
                    acc = torch.zeros(D, dtype=torch.float32)
                    for e in range(e0, e1):                       # position order, sequential fp32 adds
                        r = int(recv[s, rows_base + e])
                        if 0 <= r < local_rows:
                            acc = acc + tab[r]
                        if pos_src is not None:
                            pos_src[s * cap + e] = (s * block_floats + vec_base + b * D) // 4
                    out[s, vec_base + b * D: vec_base + (b + 1) * D] = acc
            else:
                for e in range(cap):
                    r = int(recv[s, rows_base + e])
                    if pos_src is not None:
                        pos_src[s * cap + e] = (s * block_floats + vec_base + e * D) // 4
                    if 0 <= r < local_rows:
                        out[s, vec_base + e * D: vec_base + (e + 1) * D] = tab[r]

    @staticmethod
    def combine(recv_vec, block_floats, vec_base, world, ids, pad, vocab, mode, send, block_ints, off_base, cap, n_pad,
                pad_row, dim, out):
        B, L = ids.shape
        for b in range(B):
            if L > 1:
                acc = torch.zeros(dim, dtype=torch.float32)
                for w in range(world):
                    acc = acc + recv_vec[w, vec_base + b * dim: vec_base + (b + 1) * dim]
                np_ = int(n_pad[b]) if n_pad is not None else 0
                if np_ > 0 and pad_row is not None:
                    acc = acc + float(np_) * pad_row
                if mode == POOL_MEAN:
                    acc = acc * torch.tensor(1.0 / L, dtype=torch.float32)
                out[b] = acc
            else:
                i = int(ids[b, 0])
                if pad is not None and i == pad:
                    out[b] = pad_row if pad_row is not None else 0.0
                elif 0 <= i < vocab:
                    w = i % world
                    slot = int(send[w, off_base + b])
                    out[b] = recv_vec[w, vec_base + slot * dim: vec_base + (slot + 1) * dim] if slot < cap else 0.0
                else:
                    out[b] = 0.0

    @staticmethod
    def grad_pack(grad, mode, dim, world, ids, pad, vocab, send, block_ints, off_base, cap, send_vec, block_floats, vec_base):
        B, L = ids.shape
        scale = torch.tensor(1.0 / L if (mode == POOL_MEAN and L > 1) else 1.0, dtype=torch.float32)
        for b in range(B):
            if L > 1:
                g = grad[b] * scale
                for w in range(world):
                    send_vec[w, vec_base + b * dim: vec_base + (b + 1) * dim] = g
            else:
                i = int(ids[b, 0])
                if (pad is not None and i == pad) or i < 0 or i >= vocab:
                    continue
                w = i % world
                slot = int(send[w, off_base + b])
                if slot < cap:
                    send_vec[w, vec_base + slot * dim: vec_base + (slot + 1) * dim] = grad[b]

    @staticmethod
    def segment_grad_lists(recv, world, block_ints, rows_base, cap, pos_src, local_rows, grad, piece_rows, block_floats,
                           vec_base, dim, rows_out, row_grad, n_unique, sq_norm, ws):
        acc = {}
        for s in range(world):
            for e in range(cap):
                p = s * cap + e
                r = int(recv[s, rows_base + e])
                if r < 0 or r >= local_rows:
                    continue
                if pos_src is not None:
                    o = 4 * int(pos_src[p])
                    g = grad.reshape(-1)[o:o + dim]
                else:
                    piece, i = divmod(p, piece_rows)
                    g = grad[piece, vec_base + i * dim: vec_base + (i + 1) * dim]
                acc[r] = g.clone() if r not in acc else acc[r] + g      # ascending position order
        keys = sorted(acc)
        for k, r in enumerate(keys):
            rows_out[k] = r
            row_grad[k] = acc[r]
        n_unique[0] = len(keys)
        if sq_norm is not None and keys:
            sq_norm += torch.stack([acc[r] for r in keys]).double().pow(2).sum().float()

    @staticmethod
    def segment_ws_bytes(n_pos, dim):
        return 256

    @staticmethod
    def adam(table, m, v, rows, row_grad, n_unique, coef, lr, b1, b2, eps, step_dev, lr_dev=None):
        U = int(n_unique[0])
        if U == 0:
            return
        t = float(step_dev[0])
        r = rows[:U]
        g = row_grad[:U] * (coef if coef is not None else 1.0)
        m[r] = b1 * m[r] + (1 - b1) * g
        v[r] = b2 * v[r] + (1 - b2) * g * g
        step_size = lr / (1 - b1 ** t)
        denom = v[r].sqrt() / (1 - b2 ** t) ** 0.5 + eps
        table[r] = (table[r].float() - step_size * m[r] / denom).to(table.dtype)
