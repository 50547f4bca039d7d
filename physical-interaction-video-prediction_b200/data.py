"""Synthetic push-style sequences in the on-disk layout the reference trains from (make_dataset.py:104-136 writes, per
sequence, ``image (T,64,64,3) float32 in [0,1]``, ``action (T,5)``, ``state (T,5)``; train_model.py:812-834 loads them).
The real Google push dataset cannot be downloaded here, so benchmarks and examples use this generator."""
import numpy as np


def synthetic_sequences(batch, seq_len, height=64, width=64, seed=1234):
    """A few coloured Gaussian blobs moving on a flat background plus pixel noise (so the CDNA kernels have motion to model)."""
    rs = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float32)
    seqs = []
    for _ in range(batch):
        img = np.zeros((seq_len, height, width, 3), np.float32) + rs.rand(1, 1, 1, 3).astype(np.float32) * 0.3
        for _b in range(3):
            p = rs.rand(2) * (height, width)
            vel = rs.uniform(-2, 2, 2)
            col = rs.rand(3).astype(np.float32)
            sig = rs.uniform(2, 6)
            for t in range(seq_len):
                c = p + vel * t
                img[t] += np.exp(-((yy - c[0]) ** 2 + (xx - c[1]) ** 2) / (2 * sig * sig))[..., None] * col
        img += rs.rand(seq_len, height, width, 3).astype(np.float32) * 0.05
        seqs.append([np.clip(img, 0, 1).astype(np.float32), rs.uniform(-1, 1, (seq_len, 5)).astype(np.float32),
                     rs.uniform(-1, 1, (seq_len, 5)).astype(np.float32)])
    return seqs
