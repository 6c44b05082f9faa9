"""CPU tests of the host-side mirror of the reference interface: dataset alignment, collates, label
rules, band-bin arithmetic, constructor signatures / state_dict keys / parameter counts."""
import json

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from multimodal_eeg_fmri_b200 import bridge_utils, eeg_data_utils, modules, run_training_lite
from oracle import spectral as osp

STRUCT = json.loads((GOLDEN / "structure.json").read_text())


def _keys(m):
    return {k: list(v.shape) for k, v in m.state_dict().items()}


@pytest.mark.parametrize("name,ctor", [
    ("bridge_default", lambda: modules.EEGfMRIBridgeFusionNet()),
    ("erp_v4_64_128", lambda: modules.EnhancedERPEncoder(64, 128, 2, 4)),
    ("power_v4_64_128", lambda: modules.EnhancedPowerEncoder(64, 128, 2, 4)),
    ("lite_erp_64_96", lambda: modules.LiteERPEncoder(64, 96)),
    ("lite_pw_64_96", lambda: modules.LitePowerEncoder(64, 96)),
    ("trimodal_lite_64_64_6048", lambda: modules.EnhancedTriModalFusionNetV4Lite(64, 64, 6048)),
    ("fmri_400_40000", lambda: modules.fMRIFusionNet(400, 40000)),
])
def test_state_dict_keys_and_param_counts_match_reference(name, ctor):
    m = ctor()
    assert _keys(m) == STRUCT[name]["keys"]
    assert sum(p.numel() for p in m.parameters()) == STRUCT[name]["params"]


def test_same_seed_gives_reference_initialisation():
    """Containers are built in the reference's construction order, so torch.manual_seed(s) yields the
    reference's initial weights (checked against the golden state_dict of the small bridge)."""
    torch.manual_seed(48)
    m = modules.EEGfMRIBridgeFusionNet(32, 16, 32, 2, 4, 0.0)
    z = np.load(GOLDEN / "bridge_small.npz")
    for k, v in m.state_dict().items():
        assert np.array_equal(v.numpy(), z["sd/" + k]), k


def test_fusion_weight_getters():
    w = modules.EEGfMRIBridgeFusionNet().get_fusion_weights()
    ref = STRUCT["bridge_fusion_weights_init"]
    assert w.keys() == ref.keys() and all(abs(w[k] - ref[k]) < 1e-7 for k in w)
    w = modules.fMRIFusionNet(8, 8).get_fusion_weights()
    assert w == pytest.approx(STRUCT["fmri_fusion_weights_init"])


def test_bridge_dataset_alignment_is_exact():
    eeg = {"001": torch.zeros(2), 2: torch.ones(2), "7": torch.full((2,), 7.0), 9: torch.zeros(2)}
    fmri = {1: torch.zeros(3), "2": torch.ones(3), 7: torch.full((3,), 7.0)}
    labels = {"1": 0, 2: 1, 7: 1, 9: 0}
    # sorted(subject_list) sorts the ORIGINAL keys: strings sort lexicographically ("001" < "2" < "5" < "7" < "9")
    ds = bridge_utils.BridgeFeatureDataset(eeg, fmri, labels, ["7", "2", "001", "9", "5"])
    assert [s["subject"] for s in ds.samples] == [1, 2, 7]
    with pytest.raises(TypeError):  # mixed str/int subject lists fail in sorted(), as in the reference
        bridge_utils.BridgeFeatureDataset(eeg, fmri, labels, ["7", 2])
    ds = bridge_utils.BridgeFeatureDataset(eeg, fmri, labels, [7, 2, 1, 9, 5])
    assert [s["subject"] for s in ds.samples] == [1, 2, 7]
    e, f, y, subj = bridge_utils.collate_bridge([ds[i] for i in range(len(ds))])
    assert e.shape == (3, 2) and f.shape == (3, 3) and y.dtype == torch.long and y.tolist() == [0, 1, 1] and subj == [1, 2, 7]
    assert len(bridge_utils.BridgeFeatureDataset({}, {}, {}, [1, 2])) == 0


def test_collate_balanced_dict_and_tuple():
    s = lambda i: {"erp": torch.full((2, 3), float(i)), "pw": torch.zeros(2, 3), "conn": torch.zeros(4), "label": i % 2, "subject": 10 + i}
    erp, pw, conn, y, subj = run_training_lite.collate_balanced([s(0), s(1), s(2)])
    assert erp.shape == (3, 2, 3) and y.tolist() == [0, 1, 0] and y.dtype == torch.long and subj == [10, 11, 12]
    t = lambda i: (torch.zeros(2, 3), torch.zeros(2, 3), torch.zeros(4), i, i)
    assert run_training_lite.collate_balanced([t(1), t(0)])[3].tolist() == [1, 0]


def test_load_eeg_labels(tmp_path):
    (tmp_path / "medical_score.csv").write_text(
        "Subject,Postoperative evaluation\nsub01,1\nsub02,2\nsub03,3\nsub10,5\nsub11,\n")
    assert eeg_data_utils.load_eeg_labels(tmp_path) == {1: 0, 2: 0, 3: 1, 10: 1}
    assert eeg_data_utils.load_eeg_labels(tmp_path, binary=False) == {1: 0, 2: 0, 3: 3, 10: 5}
    with pytest.raises(FileNotFoundError):
        eeg_data_utils.load_eeg_labels(tmp_path / "nope")


@pytest.mark.parametrize("nfft,fs", [(1024, 1000.0), (512, 250.0), (128, 128.0), (2048, 1000.0), (512, 512.0)])
def test_band_bins_match_oracle(nfft, fs):
    bands = list(eeg_data_utils.DEFAULT_BANDS.values())
    assert eeg_data_utils.band_bins(bands, nfft, fs) == osp.band_bins(bands, nfft, fs).tolist()


def test_aggregate_rejects_unknown_method():
    from multimodal_eeg_fmri_b200 import fmri_utils
    with pytest.raises(ValueError):
        fmri_utils.aggregate_roi_timeseries(torch.zeros(1, 2, 3), "median")


def test_dropout_seed_stream_can_be_saved_and_restored():
    """functional.seed_state / set_seed_state: the host seeds a CUDA-graph capture freezes can be re-drawn (the eager twin
    of a replay, tests/test_gpu_graphed_step.py) and a warm-up step leaves the stream where it was."""
    from multimodal_eeg_fmri_b200 import functional as XF
    XF.manual_seed(1234)
    first = [XF.next_seed() for _ in range(3)]
    state = XF.seed_state()
    later = [XF.next_seed() for _ in range(4)]
    XF.set_seed_state(state)
    assert [XF.next_seed() for _ in range(4)] == later
    XF.manual_seed(1234)
    assert [XF.next_seed() for _ in range(3)] == first
    assert len(set(first + later)) == 7 and all(0 < s < 2 ** 63 for s in first + later)


def test_device_floats_reads_on_access_and_snapshots():
    """modules.DeviceFloats: the fusion-weight dict of the lite net (crossmodal_v4_enhancements.py:803-809 returns floats)
    reads its tensor when looked at; dict(w) / copy / pickle give plain snapshots."""
    import copy
    import json
    import pickle
    from multimodal_eeg_fmri_b200.modules import DeviceFloats
    t = torch.tensor([0.25, 0.5, 1.25])
    w = DeviceFloats(("erp_weight", "pw_weight", "conn_weight"), t)
    assert w["erp_weight"] == 0.25 and w.get("conn_weight") == 1.25 and w.get("missing", 7) == 7
    assert list(w) == ["erp_weight", "pw_weight", "conn_weight"] and len(w) == 3 and "pw_weight" in w
    snap = dict(w)
    assert type(snap) is dict and snap == {"erp_weight": 0.25, "pw_weight": 0.5, "conn_weight": 1.25} and w == snap
    t[0] = 0.75  # what a graph replay does to the static tensor
    assert w["erp_weight"] == 0.75 and snap["erp_weight"] == 0.25 and w != snap
    assert {**w}["erp_weight"] == 0.75 and w.copy() == dict(w.items())
    assert type(copy.deepcopy(w)) is dict and pickle.loads(pickle.dumps(w)) == dict(w)
    assert json.loads(json.dumps(dict(w))) == dict(w) and all(isinstance(v, float) for v in w.values())
