"""vae_b200.data against the reference's prepare.py semantics on small CSV fixtures written here."""
import os

import numpy as np
import pandas as pd
import pytest

from vae_b200 import data


def _write(tmp_path, with_folds):
    df = pd.DataFrame({"user": [10, 10, 30, 20, 30, 20, 10, 40], "item": [7, 5, 5, 9, 7, 5, 9, 7],
                       "rating": [5, 3, 4, 1, 2, 5, 4, 3]})
    df["outcome"] = (df["rating"] >= 4).astype(int)                  # prepare.py:54
    df.to_csv(tmp_path / "data.csv", index=False)
    if with_folds:
        pd.DataFrame({"index": [0, 1, 2, 3, 4, 5]}).to_csv(tmp_path / "trainval.csv", index=False)
        pd.DataFrame({"index": [6, 7]}).to_csv(tmp_path / "test.csv", index=False)
    return df


def test_load_data_with_fold_files_matches_prepare_py(tmp_path):
    _write(tmp_path, True)
    d = data.load_data(str(tmp_path), "reg")
    assert (d.n_users, d.n_items) == (4, 3)
    # users 10,20,30,40 -> 0..3; items 5,7,9 -> 0..2, shifted by n_users (prepare.py:45-47)
    assert d.x_train.tolist() == [[0, 5], [0, 4], [2, 4], [1, 6], [2, 5], [1, 4]]
    assert d.x_test.tolist() == [[0, 6], [3, 5]]
    assert d.y_train.tolist() == [5, 3, 4, 1, 2, 5] and d.y_test.tolist() == [4, 3]
    tc = d.train_counts()
    assert tc.tolist() == [2, 2, 2, 1, 3, 2, 1]          # user 3 (id 40) unseen in training -> 1 (N9)
    assert data.load_data(str(tmp_path), "class").y_train.tolist() == [1, 0, 1, 0, 0, 1]


def test_seeded_split_without_fold_files_and_batches(tmp_path):
    _write(tmp_path, False)
    a, b = data.load_data(str(tmp_path), "class"), data.load_data(str(tmp_path), "class")
    assert np.array_equal(a.folds["test"], b.folds["test"]) and len(a.y_test) == 2 and len(a.y_train) == 6
    assert sorted(a.folds["test"].tolist() + a.folds["trainval"].tolist()) == list(range(8))
    got = list(data.batches(a.x_train, a.y_train, 4))
    assert [len(x) for x, _ in got] == [4, 2]             # contiguous, in order, last one short (N7/N8)
    assert np.array_equal(np.concatenate([x.numpy() for x, _ in got]), a.x_train)
    with pytest.raises(KeyError):
        (tmp_path / "data.csv").write_text("user,item,outcome\n0,0,1\n")
        data.load_data(str(tmp_path), "reg")


def test_libfm_export_format(tmp_path):
    data.write_libfm(str(tmp_path / "t.libfm"), np.array([[0, 5], [2, 4]]), np.array([1, 0]))
    assert (tmp_path / "t.libfm").read_text() == "1 0:1 5:1\n0 2:1 4:1\n"


@pytest.mark.skipif(not os.path.isfile("/root/reference/data/fraction/data.csv"), reason="reference not mounted")
def test_bundled_fraction_dataset_shape():
    d = data.load_data("/root/reference/data/fraction", "class")
    assert (d.n_users, d.n_items) == (536, 20) and len(d.y_train) + len(d.y_test) == 10720
    assert d.x_train[:, 1].min() >= 536 and d.train_counts().shape == (556,)
