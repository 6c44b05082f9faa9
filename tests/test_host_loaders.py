"""On-disk formats either side of the path against what the REFERENCE loaders
return on the committed fixture tree tests/golden/data/ (oracle/make_golden.py `loaders` wrote both the tree and
tests/golden/loaders.npz by running fmri_utils.py:115-241, eeg_data_utils.py:19-186 and the BridgeRawDataset
source of _test_bridge.py:391-453).  Byte / integer work is bit-exact; the ROI mean/std aggregation (the only
arithmetic; a device op) is held to 1e-5.  The CPU tests run that op through tests/fake_ops.py; the `gpu` test
runs the real kernel."""
import logging

import numpy as np
import pytest
import torch

import fake_ops
from conftest import GOLDEN, assert_close_rel
from multimodal_eeg_fmri_b200 import bridge_utils as bu
from multimodal_eeg_fmri_b200 import eeg_data_utils as edu
from multimodal_eeg_fmri_b200 import fmri_utils as fu

DATA = GOLDEN / "data"
FM = DATA / "fmri"
EEG = DATA / "eeg"
SUBJECTS = [1, 2, 3, 5]
BANDS = {"alpha": "Alpha", "beta": "Beta"}


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLDEN / "loaders.npz")


def _group(z, prefix):
    out = {}
    for k in z.files:
        head, _, rest = k.partition("/")
        if head == prefix and not rest.count("/"):
            out[rest] = z[k]
    return out


def _key(k):
    return "|".join(map(str, k)) if isinstance(k, tuple) else str(k)


def _same_dict(ours, ref, exact=True, what=""):
    assert sorted(_key(k) for k in ours) == sorted(ref), f"{what}: keys {sorted(map(_key, ours))} != {sorted(ref)}"
    for k, v in ours.items():
        a, b = np.asarray(v), ref[_key(k)]
        assert a.shape == b.shape and a.dtype == b.dtype, f"{what}[{k}]: {a.dtype}{a.shape} vs {b.dtype}{b.shape}"
        if exact:
            assert np.array_equal(a, b), f"{what}[{k}] differs"
        else:
            assert_close_rel(a, b, 1e-5, f"{what}[{k}]")


def check_activation_loader(gold, device):
    for agg in ("both", "mean", "std"):
        got = fu.load_activation_features(FM, SUBJECTS, ["taskA", "taskB"], agg, device=device)
        assert all(v.dtype == torch.float32 and v.device.type == "cpu" for v in got.values())
        assert list(got) == [1, 2, 3]  # subject_list order; 5 has only an unreadable file, 4 no directory
        _same_dict(got, _group(gold, f"act_{agg}"), exact=False, what=f"activation/{agg}")
    got = fu.load_activation_features(FM, SUBJECTS, ["taskB"], "both", device=device)
    _same_dict(got, _group(gold, "act_both_B_only"), exact=False, what="activation/B only")


def test_activation_features_vs_reference(gold, monkeypatch, caplog):
    fake_ops.install(monkeypatch)
    with caplog.at_level(logging.WARNING):
        check_activation_loader(gold, "cpu")
    assert any("subject_5_activation_taskA.csv" in r.getMessage() for r in caplog.records)  # logged, not raised
    # an unknown aggregation is logged per file and yields {} -- the reference raises its ValueError inside its own try
    assert _group(gold, "act_median") == {}
    assert fu.load_activation_features(FM, SUBJECTS, ["taskA", "taskB"], "median", device="cpu") == {}
    assert fu.load_activation_features(FM, [], ["taskA"], device="cpu") == {}
    twice = fu.load_activation_features(FM, [2, 1, 2], ["taskA", "taskB"], device="cpu")  # repeated subject: one entry
    assert list(twice) == [2, 1]
    _same_dict({2: twice[2]}, {"2": gold["act_both/2"]}, exact=False, what="repeated subject")
    with pytest.raises(ValueError):
        fu.aggregate_roi_timeseries(torch.zeros(1, 2, 3), "median")


def test_activation_loader_has_no_cpu_fallback():
    """Without the fake backend the aggregation op refuses host tensors: the product path needs the CUDA library."""
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    from multimodal_eeg_fmri_b200._lib import XmodalError
    with pytest.raises(XmodalError):
        fu.load_activation_features(FM, [1], ["taskA"])


@pytest.mark.gpu
def test_activation_features_vs_reference_gpu(gold):
    check_activation_loader(gold, "cuda")


def test_connectivity_and_label_loaders_bit_exact(gold):
    got = fu.load_connectivity_features(FM, SUBJECTS, ["rest", "task"])
    _same_dict(got, _group(gold, "conn"), what="connectivity")
    assert got[2].numel() == 16 and got[1].numel() == 32  # subject 2 lacks the `task` matrix
    for name, path, subjects in (("labels_str", DATA / "labels_str", [1, 2, 3, 5]), ("labels_num", DATA / "labels_num" / "inner", [1, 2, 3])):
        ref = {int(k): int(v) for k, v in _group(gold, name).items()}
        assert fu.load_fmri_labels(path, subjects) == ref
    with pytest.raises(ValueError) as e:
        fu.load_fmri_labels(DATA / "labels_bad", [1])
    assert str(e.value).replace(str(DATA / "labels_bad"), "<dir>") == str(gold["labels_bad_error"])
    dummy = fu.load_fmri_labels(DATA / "eeg" / "conn", [4, 8])  # no label file anywhere: random 0/1 per subject
    assert set(dummy) == {4, 8} and set(dummy.values()) <= {0, 1}


def test_eeg_mat_loaders_bit_exact(gold, caplog):
    conn = edu.load_eeg_conn_features(EEG / "conn", [1, 2, 3], BANDS, ["open", "close"])
    pw = edu.load_eeg_pw_features(EEG / "pw", [1, 2, 3], ["alpha", "beta"], ["1_Hz", "2_Hz"])
    with caplog.at_level(logging.WARNING):
        erp = edu.load_eeg_erp_features(EEG / "erp", [1, 2, 3], ["alpha", "beta"], ["1_Hz", "2_Hz"])
    assert any("ERP_sub02_beta_2_Hz.mat" in r.getMessage() for r in caplog.records)
    _same_dict(conn, _group(gold, "eeg_conn"), what="eeg conn")
    _same_dict(pw, _group(gold, "eeg_pw"), what="eeg pw")
    _same_dict(erp, _group(gold, "eeg_erp"), what="eeg erp")
    assert erp[(1, "alpha", "1_Hz", 0)].shape == (3, 5) and erp[(1, "alpha", "1_Hz", 0)][0, 0] == 0.0  # NaN -> 0, shape kept
    for name, binary in (("eeg_labels_binary", True), ("eeg_labels_raw", False)):
        ref = {int(k): v.item() for k, v in _group(gold, name).items()}
        got = edu.load_eeg_labels(EEG / "labels", binary)
        assert got == ref and all(type(got[k]) is type(ref[k]) or float(got[k]) == float(ref[k]) for k in got)
    with pytest.raises(FileNotFoundError):
        edu.load_eeg_labels(EEG / "conn")


def test_bridge_raw_dataset_alignment_bit_exact(gold, monkeypatch):
    fake_ops.install(monkeypatch)
    conn = edu.load_eeg_conn_features(EEG / "conn", [1, 2, 3], BANDS, ["open", "close"])
    pw = edu.load_eeg_pw_features(EEG / "pw", [1, 2, 3], ["alpha", "beta"], ["1_Hz", "2_Hz"])
    erp = edu.load_eeg_erp_features(EEG / "erp", [1, 2, 3], ["alpha", "beta"], ["1_Hz", "2_Hz"])
    f_act = fu.load_activation_features(FM, SUBJECTS, ["taskA", "taskB"], "both", device="cpu")
    f_conn = fu.load_connectivity_features(FM, SUBJECTS, ["rest", "task"])
    labels = edu.load_eeg_labels(EEG / "labels", True)
    labels[3] = 1
    ds = bu.BridgeRawDataset(erp, pw, conn, f_act, f_conn, labels, ["3", "2", "1", "5", "4", "10"], BANDS, ["open", "close"])
    assert [ds[i][4] for i in range(len(ds))] == gold["raw_ds/subjects"].tolist()
    assert [ds[i][3] for i in range(len(ds))] == gold["raw_ds/labels"].tolist()
    assert [len(ds[i][0]) for i in range(len(ds))] == gold["raw_ds/n_eeg"].tolist()
    for i in range(len(ds)):
        for j, (e, p, c) in enumerate(ds[i][0]):  # subject 3 has no power file: zero-padded like the first entry
            for name, a in (("erp", e), ("pw", p), ("conn", c)):
                b = gold[f"raw_ds/{i}/{j}/{name}"]
                assert a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b), (i, j, name)
        assert_close_rel(ds[i][1], gold[f"raw_ds/{i}/fmri_act"], 1e-5, "fmri_act")
        assert np.array_equal(ds[i][2].numpy(), gold[f"raw_ds/{i}/fmri_conn"])
    # zero stand-ins for missing entries are fresh arrays (reference :416-421): writing to one leaves the others alone
    one = np.ones((2, 3), np.float32)
    erp2 = {(1, "alpha", "1_Hz", "a"): one, (1, "beta", "1_Hz", "a"): one, (1, "beta", "2_Hz", "a"): one}
    pw2 = {(1, "alpha", "1_Hz", "a"): one}
    ds2 = bu.BridgeRawDataset(erp2, pw2, {}, {1: one}, {1: one}, {1: 0}, [1], BANDS, ["open"])
    assert len(ds2) == 0  # no connectivity dict at all: nothing to shape the stand-in after, entries dropped
    conn2 = {(1, "alpha", "open", "a"): one}
    ds2 = bu.BridgeRawDataset(erp2, pw2, conn2, {1: one}, {1: one}, {1: 0}, [1], BANDS, ["open"])
    pads = [p for (_, p, _) in ds2[0][0] if not p.any()] + [c for (_, _, c) in ds2[0][0] if not c.any()]
    assert len(pads) == 4 and len({id(p) for p in pads}) == 4
    pads[0][...] = 7.0
    assert not any(p.any() for p in pads[1:])
    empty = bu.BridgeRawDataset({}, {}, {}, {}, {}, {}, [1], BANDS, ["open"])
    assert len(empty) == 0
