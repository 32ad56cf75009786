"""A numpy stand-in for mmiss_b200.index.DeviceIndex, used ONLY by the CPU tests of the host-side
logic (Collection bookkeeping, sharded exchange).  It answers queries with the oracle; it is test
infrastructure and is never importable from the product package."""
import numpy as np

from oracle import cosine_oracle as O


class FakeIndex:
    def __init__(self, dim, dtype="f32", device=0, capacity=0, row_base=0, row_stride=1):
        self.dim, self.dtype, self.row_base = dim, ("bf16" if dtype in ("bf16", "bfloat16") else "f32"), row_base
        self.row_stride = row_stride
        self.X = np.zeros((0, dim), np.float32)
        self.bits = []
        self.closed = False

    def __len__(self):
        return self.X.shape[0]

    def add(self, rows):
        rows = np.asarray(rows, np.float32).reshape(-1, self.dim)
        first = len(self)
        self.X = np.concatenate([self.X, rows])
        self.bits += [set() for _ in range(rows.shape[0])]
        return first

    def remove(self, row):
        last = len(self) - 1
        moved = -1 if row == last else last
        if row != last:
            self.X[row] = self.X[last]
            self.bits[row] = self.bits[last]
        self.X = self.X[:last]
        self.bits.pop()
        return moved

    def set_row(self, row, vec):
        self.X[row] = np.asarray(vec, np.float32).reshape(-1)

    def clear(self):
        self.X = self.X[:0]
        self.bits = []

    def set_filter_bits_range(self, first, bits_lists):
        for j, bits in enumerate(bits_lists):
            self.bits[first + j] = set(bits)

    def set_filter_bits(self, row, bits):
        self.bits[row] = set(bits)

    def get_filter_bits(self, row):
        return sorted(self.bits[row])

    def get_rows(self, first, n):
        x = self.X[first:first + n]
        return O.bf16_round(x) if self.dtype == "bf16" else x.copy()

    def query(self, q, k, require_bits=None, mode="auto"):
        q = np.asarray(q, np.float32).reshape(-1, self.dim)
        valid = None
        if require_bits is not None:
            valid = np.array([set(require_bits) <= b for b in self.bits], bool)
        s, r = O.cosine_topk(q, self.X, k, corpus_dtype=self.dtype, valid=valid)
        out_s = np.full((q.shape[0], k), -np.inf, np.float32)
        out_r = np.full((q.shape[0], k), -1, np.int64)
        out_s[:, :s.shape[1]], out_r[:, :r.shape[1]] = s, r * self.row_stride + self.row_base
        return out_s, out_r

    def query_multimodal(self, img, txt, w, k, require_bits=None, mode="auto"):
        img = np.asarray(img, np.float32).reshape(-1, self.dim)
        txt = np.asarray(txt, np.float32).reshape(-1, self.dim)
        w = np.broadcast_to(np.asarray(w, np.float64), (img.shape[0],))
        q = np.stack([O.blend(img[i], txt[i], float(w[i])) for i in range(img.shape[0])])
        return self.query(q, k, require_bits, mode)

    def blend_dev(self, img, txt, w, out=None, stream=None):
        import torch
        a, t, ww = img.cpu().numpy(), txt.cpu().numpy(), w.cpu().numpy()
        return torch.from_numpy(np.stack([O.blend(a[i], t[i], float(ww[i])) for i in range(a.shape[0])]).astype(np.float32))

    def apply_filter_sweep(self, prompt, tau, bit):
        m = O.filter_mask(np.asarray(prompt, np.float32).reshape(1, self.dim), self.X, tau)[0]
        for r, hit in enumerate(m.tolist()):
            self.bits[r] = (self.bits[r] - {bit}) | ({bit} if hit else set())
        return int(m.sum())

    def filter_words(self):
        return (len(self) + 255) // 256 * 8

    def filter_sweep(self, prompts, tau):
        m = O.filter_mask(np.asarray(prompts, np.float32).reshape(-1, self.dim), self.X, tau)
        packed = O.pack_mask_bits(m)
        out = np.zeros((m.shape[0], self.filter_words()), np.uint32)
        out[:, :packed.shape[1]] = packed
        return out

    def dedup(self, tau, row_lo=0, row_hi=None, capacity=0):
        i, j, s = O.dedup_pairs(self.X, tau)
        hi = len(self) if row_hi is None else row_hi
        keep = (i >= row_lo) & (i < hi)
        return i[keep], j[keep], s[keep]

    def close(self):
        self.closed = True
