"""CPU: the seam oracle (oracle/seam_oracle.py, knn_oracle.derived_feature_types) against the goldens that
oracle/gen_golden.py produced by executing the reference's own source lines."""
import numpy as np

from oracle import knn_oracle as ko
from oracle import seam_oracle as so
from tests import _golden


def test_softmax_mean_oracle_matches_reference_golden():
    for name in ("seam_softmax_mean_n3_t133_c96", "seam_softmax_mean_n2_t7_c45"):
        g = _golden.load(name)
        out = so.softmax_mean(g.inputs["logits"])
        np.testing.assert_allclose(out, g.outputs["out"], rtol=2e-6, atol=1e-9)
        dx = so.softmax_mean_backward(g.inputs["logits"], g.inputs["grad_out"])
        # the golden is torch float32 autograd: (g - sum g p) cancels, so the tolerance is relative to the largest entry
        ref = g.outputs["grad_logits"]
        np.testing.assert_allclose(dx, ref, rtol=2e-5, atol=2e-6 * np.abs(ref).max())


def test_embed_handoff_oracle_matches_reference_golden():
    g = _golden.load("seam_embed_handoff_n8_d768")
    got = np.concatenate([so.f_normalize_rows(g.inputs["batch0"]), so.f_normalize_rows(g.inputs["batch1"])])
    ref = g.outputs["features"]
    assert got.dtype == np.float64 and ref.dtype == np.float64
    # float32 values widened to float64; torch's reduction order may differ in the last float32 bit
    np.testing.assert_allclose(got, ref, rtol=3e-7, atol=0)
    assert np.all(ref[6] == 0) and np.all(got[6] == 0)  # the zero row stays zero (eps clamp)


def test_derived_feature_types_oracle_matches_reference_golden():
    g = _golden.load("seam_derived_feature_types_n6_d16")
    img, dna, txt = g.inputs["image"], g.inputs["dna"], g.inputs["text"]
    labels = [{"species": f"s{i}"} for i in range(6)]
    d = ko.derived_feature_types(img, dna, txt, for_key_set=True, labels=labels)
    assert np.array_equal(d["averaged_feature"], g.outputs["averaged_feature"])
    assert np.array_equal(d["concatenated_feature"], g.outputs["concatenated_feature"])
    assert np.array_equal(d["all_key_features"], g.outputs["all_key_features"])
    assert [l["species"] for l in d["all_key_features_label"]] == list(g.outputs["all_key_features_label_species"])
    d2 = ko.derived_feature_types(img, dna, txt, for_key_set=False, labels=labels)
    assert d2["all_key_features"] is None and d2["all_key_features_label"] is None
