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


# ----------------------------------------------------------------------------- on-disk data set (train_model.py:812-843)
MAP_HEADER = ['id', 'img_bitmap_path', 'img_np_path', 'action_np_path', 'state_np_path', 'img_bitmap_pred_path', 'img_np_pred_path']


def write_dataset(data_dir, seqs):
    """Write sequences in the layout ``make_dataset.py:130-156`` produces: per sequence j ``image_batch_j.npy (T,H,W,3)``,
    ``action_batch_j.npy (T,5)``, ``state_batch_j.npy (T,5)`` plus ``map.csv`` (all fields quoted, header row first)."""
    import csv
    import os
    os.makedirs(data_dir, exist_ok=True)
    rows = []
    for j, (img, act, sta) in enumerate(seqs):
        np.save(os.path.join(data_dir, "image_batch_%d" % j), np.asarray(img, np.float32))
        np.save(os.path.join(data_dir, "action_batch_%d" % j), np.asarray(act, np.float32))
        np.save(os.path.join(data_dir, "state_batch_%d" % j), np.asarray(sta, np.float32))
        rows.append([j, "", "image_batch_%d.npy" % j, "action_batch_%d.npy" % j, "state_batch_%d.npy" % j, "", ""])
    with open(os.path.join(data_dir, "map.csv"), "w", newline="") as f:
        w = csv.writer(f, quoting=csv.QUOTE_ALL)
        w.writerow(MAP_HEADER)
        for r in rows:
            w.writerow(r)


def load_dataset(data_dir):
    """train_model.py:812-834: read ``map.csv`` (first row is the header), load every sequence into RAM as float32.
    Returns (images (N,T,H,W,3), actions (N,T,5), states (N,T,5)).  Raises like the reference exits when the map is empty."""
    import csv
    import os
    with open(os.path.join(data_dir, "map.csv"), "r", newline="") as f:
        data_map = [row for row in csv.reader(f)]
    if len(data_map) <= 1:                                   # empty or only header (ref:819-821)
        raise ValueError("No file map found")
    images, actions, states = [], [], []
    for row in data_map[1:]:
        images.append(np.float32(np.load(os.path.join(data_dir, row[2]))))
        actions.append(np.float32(np.load(os.path.join(data_dir, row[3]))))
        states.append(np.float32(np.load(os.path.join(data_dir, row[4]))))
    return np.asarray(images, np.float32), np.asarray(actions, np.float32), np.asarray(states, np.float32)


def train_val_split(images, actions, states, split=0.95):
    """train_model.py:836-843: the first floor(split * N) sequences train, the rest validate.  Returns two lists of [img, act, sta] groups
    (train_model.py:896-911)."""
    k = int(np.floor(split * len(images)))
    group = lambda lo, hi: [[images[i], actions[i], states[i]] for i in range(lo, hi)]
    return group(0, k), group(k, len(images))


class SerialIterator(object):
    """``chainer.iterators.SerialIterator(dataset, batch_size, repeat, shuffle)`` as train_model.py:914-915 uses it (Chainer 2.0.1
    semantics): one ``numpy.random.permutation`` at construction, batches walk that order, the order is re-shuffled IN PLACE
    (``numpy.random.shuffle``) when an epoch ends and the last batch of an epoch is topped up from the new order.  It draws from the
    legacy GLOBAL NumPy stream -- the same stream ``scheduled_sample`` shuffles with (train_model.py:93-96) -- so a seeded run
    interleaves the two exactly like the reference."""

    def __init__(self, dataset, batch_size, repeat=True, shuffle=True):
        self.dataset, self.batch_size, self._repeat = dataset, int(batch_size), repeat
        self._order = np.random.permutation(len(dataset)) if shuffle else None
        self._shuffle = shuffle
        self.reset_position()

    def reset_position(self):
        self.current_position = 0
        self.epoch = 0
        self.is_new_epoch = False

    def reset(self):
        if self._shuffle:
            self._order = np.random.permutation(len(self.dataset))
        self.reset_position()

    def __iter__(self):
        return self

    def __next__(self):
        if not self._repeat and self.epoch > 0:
            raise StopIteration
        i, N = self.current_position, len(self.dataset)
        i_end = i + self.batch_size
        pick = (lambda lo, hi: list(self.dataset[lo:hi])) if self._order is None else (lambda lo, hi: [self.dataset[k] for k in self._order[lo:hi]])
        batch = pick(i, i_end)
        if i_end >= N:
            if self._repeat:
                rest = i_end - N
                if self._order is not None:
                    np.random.shuffle(self._order)
                if rest > 0:
                    batch.extend(pick(0, rest))
                self.current_position = rest
            else:
                self.current_position = N
            self.epoch += 1
            self.is_new_epoch = True
        else:
            self.is_new_epoch = False
            self.current_position = i_end
        return batch

    next = __next__

    @property
    def epoch_detail(self):
        return self.epoch + self.current_position / float(len(self.dataset))
