"""CPU tests of the data side: ``map.csv`` + ``.npy`` layout of make_dataset.py:130-156, the loader of
train_model.py:812-834, the train/validation split :836-843 and the SerialIterator semantics of :914-915 (Chainer 2.0.1)."""
import csv
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.fixture(scope="module")
def D():
    import pivp_b200
    return pivp_b200.data


def test_dataset_files_follow_the_reference_layout(D, tmp_path):
    seqs = D.synthetic_sequences(5, 4, 16, 24, seed=3)
    d = str(tmp_path / "push_train")
    D.write_dataset(d, seqs)
    with open(os.path.join(d, "map.csv"), newline="") as f:
        rows = list(csv.reader(f))
    assert rows[0] == ['id', 'img_bitmap_path', 'img_np_path', 'action_np_path', 'state_np_path', 'img_bitmap_pred_path', 'img_np_pred_path']
    assert rows[3][2:5] == ["image_batch_2.npy", "action_batch_2.npy", "state_batch_2.npy"]       # columns the loader indexes (ref:829-831)
    assert open(os.path.join(d, "map.csv")).read().startswith('"id","img_bitmap_path"')           # csv.QUOTE_ALL (make_dataset.py:152)
    img, act, sta = D.load_dataset(d)
    assert img.shape == (5, 4, 16, 24, 3) and img.dtype == np.float32 and act.shape == (5, 4, 5) and sta.shape == (5, 4, 5)
    assert all(np.array_equal(img[i], seqs[i][0]) and np.array_equal(act[i], seqs[i][1]) for i in range(5))
    tr, va = D.train_val_split(img, act, sta, 0.7)                                                 # floor(0.7 * 5) = 3
    assert len(tr) == 3 and len(va) == 2 and np.array_equal(va[0][2], seqs[3][2])


def test_empty_map_is_an_error(D, tmp_path):
    d = str(tmp_path / "empty")
    os.makedirs(d)
    with open(os.path.join(d, "map.csv"), "w") as f:
        f.write('"id","img_bitmap_path","img_np_path","action_np_path","state_np_path","img_bitmap_pred_path","img_np_pred_path"\n')
    with pytest.raises(ValueError, match="No file map found"):                                    # ref:819-821
        D.load_dataset(d)


def literal_serial_iterator(n, batch_size, steps):
    """Chainer 2.0.1 SerialIterator(repeat=True, shuffle=True).__next__, written out literally on indices."""
    order = np.random.permutation(n)
    pos, epoch, out = 0, 0, []
    for _ in range(steps):
        i, i_end = pos, pos + batch_size
        batch = list(order[i:i_end])
        new_epoch = False
        if i_end >= n:
            rest = i_end - n
            np.random.shuffle(order)
            if rest > 0:
                batch.extend(order[:rest])
            pos = rest
            epoch += 1
            new_epoch = True
        else:
            pos = i_end
        out.append((batch, epoch, new_epoch, pos))
    return out


@pytest.mark.parametrize("n,bs", [(10, 4), (8, 4), (7, 7), (5, 2)])
def test_serial_iterator_consumes_the_global_rng_like_chainer(D, n, bs):
    data = list(range(100, 100 + n))
    np.random.seed(42)
    want = literal_serial_iterator(n, bs, 9)
    tail_want = np.random.rand()                       # the global stream continues identically afterwards
    np.random.seed(42)
    it = D.SerialIterator(data, bs, repeat=True, shuffle=True)
    for batch, epoch, new_epoch, pos in want:
        ep_before = it.epoch
        got = it.next()
        assert got == [data[i] for i in batch]
        assert (it.epoch, it.is_new_epoch, it.current_position) == (epoch, new_epoch, pos)
        assert len(got) == bs and it.epoch - ep_before in (0, 1)
    assert np.random.rand() == tail_want


def test_serial_iterator_without_repeat_stops_after_one_epoch(D):
    it = D.SerialIterator(list(range(5)), 2, repeat=False, shuffle=False)
    got = [b for b in it]
    assert got == [[0, 1], [2, 3], [4]] and it.epoch == 1
    it.reset()
    assert it.epoch == 0 and it.next() == [0, 1]


def test_oracle_resize_images_is_align_corners_bilinear():
    from oracle.fused_ops import resize_images
    x = np.arange(2 * 3 * 4 * 6, dtype=np.float32).reshape(2, 3, 4, 6)
    assert np.array_equal(resize_images(x, (4, 6)), x)                                 # identity size: exact
    y = resize_images(x, (7, 11))
    assert y.shape == (2, 3, 7, 11)
    assert np.allclose(y[:, :, 0, 0], x[:, :, 0, 0]) and np.allclose(y[:, :, -1, -1], x[:, :, -1, -1])   # corners aligned
    # a linear ramp is reproduced exactly by bilinear interpolation
    assert np.allclose(y[0, 0, 0], np.linspace(x[0, 0, 0, 0], x[0, 0, 0, -1], 11), atol=1e-5)
