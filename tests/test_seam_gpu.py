"""GPU parity of the seam kernels (csrc/seam.cu) through the C ABI: BarcodeBERT head forward/backward,
embedding hand-off, derived feature types feeding the eval entry point."""
import numpy as np
import pytest
import torch

from oracle import knn_oracle as ko
from oracle import seam_oracle as so
from tests import _golden

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda:0")


@pytest.mark.parametrize("name", ["seam_softmax_mean_n3_t133_c96", "seam_softmax_mean_n2_t7_c45"])
def test_softmax_mean_matches_reference_golden(name):
    from clibd_b200 import seam
    g = _golden.load(name)
    x = torch.from_numpy(g.inputs["logits"]).to(_dev()).requires_grad_(True)
    out = seam.softmax_mean(x)
    out.backward(torch.from_numpy(g.inputs["grad_out"]).to(_dev()))
    np.testing.assert_allclose(out.detach().cpu().numpy(), g.outputs["out"], rtol=1e-5, atol=1e-8)
    ref = g.outputs["grad_logits"]  # float32 autograd: (g - sum g p) cancels -> tolerance relative to the largest entry
    np.testing.assert_allclose(x.grad.cpu().numpy(), ref, rtol=1e-4, atol=4e-6 * np.abs(ref).max())


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1e-2), (torch.float16, 2e-3)])
@pytest.mark.parametrize("n,t,c", [(500, 133, 768), (7, 133, 1024), (5, 20, 2048), (9, 3, 100), (4, 11, 45), (1, 1, 8)])
def test_softmax_mean_matches_oracle(dtype, tol, n, t, c):
    from clibd_b200 import seam
    gen = torch.Generator().manual_seed(n * 1000 + c)
    logits = (torch.randn(n, t, c, generator=gen) * 4).to(dtype)
    gout = torch.randn(n, c, generator=gen).to(dtype)
    x = logits.to(_dev()).requires_grad_(True)
    out = seam.softmax_mean(x)
    assert out.dtype == dtype and out.shape == (n, c)
    out.backward(gout.to(_dev()))
    ref = so.softmax_mean(logits.float().numpy())
    rdx = so.softmax_mean_backward(logits.float().numpy(), gout.float().numpy())
    got, gdx = out.detach().float().cpu().numpy(), x.grad.float().cpu().numpy()
    # outputs are rounded to `dtype`; compare norm-wise and element-wise with the dtype's tolerance
    assert np.linalg.norm(got - ref) <= tol * np.linalg.norm(ref)
    assert np.linalg.norm(gdx - rdx) <= tol * np.linalg.norm(rdx) + 1e-30
    np.testing.assert_allclose(got, ref, rtol=4 * tol, atol=tol * np.abs(ref).max())
    # rows of probabilities sum to one -> the output sums to one per sample, the gradient to zero per token row
    assert np.allclose(got.sum(1), 1.0, atol=4 * tol)
    assert np.abs(gdx.sum(-1)).max() <= 4 * tol * np.abs(gdx).max() * c ** 0.5 + 1e-30


def test_softmax_mean_errors():
    from clibd_b200 import seam
    with pytest.raises(RuntimeError):
        seam.softmax_mean(torch.zeros(2, 3, 8))
    with pytest.raises(ValueError):
        seam.softmax_mean(torch.zeros(2, 8, device=_dev()))
    with pytest.raises(TypeError):
        seam.softmax_mean(torch.zeros(2, 3, 8, device=_dev(), dtype=torch.float64))
    assert seam.softmax_mean(torch.zeros(0, 3, 8, device=_dev())).shape == (0, 8)
    out = seam.softmax_mean(torch.full((2, 3, 8), -1e30, device=_dev()))  # uniform rows, no NaN
    assert torch.allclose(out, torch.full_like(out, 1 / 8))


def test_embedding_store_matches_reference_golden():
    from clibd_b200 import seam
    g = _golden.load("seam_embed_handoff_n8_d768")
    st = seam.EmbeddingStore()
    st.append(torch.from_numpy(g.inputs["batch0"]).to(_dev()))
    st.append(torch.from_numpy(g.inputs["batch1"]).to(_dev()))
    got = st.tensor().cpu().numpy().astype(np.float64)
    np.testing.assert_allclose(got, g.outputs["features"], rtol=3e-7, atol=0)
    assert np.all(got[6] == 0)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("d", [768, 1536, 45, 4100])
def test_embedding_store_grows_and_matches_oracle(dtype, d):
    from clibd_b200 import seam
    gen = torch.Generator().manual_seed(d)
    st = seam.EmbeddingStore(capacity=4)
    chunks = [torch.randn(n, d, generator=gen).to(dtype) for n in (3, 1500, 1, 700)]
    for c in chunks:
        st.append(c.to(_dev()))
    got = st.tensor().cpu().numpy().astype(np.float64)
    ref = np.concatenate([so.f_normalize_rows(c.float().numpy()) for c in chunks])
    assert got.shape == ref.shape == (2204, d)
    np.testing.assert_allclose(got, ref, rtol=4e-7, atol=0)
    with pytest.raises(ValueError):
        st.append(torch.zeros(2, d + 1, device=_dev()))


class _ToyModel(torch.nn.Module):
    """stands in for SimpleCLIP: three linear encoders, outputs already normalised (simple_clip.py:45-60)"""

    def __init__(self, d_in, d):
        super().__init__()
        self.enc = torch.nn.ModuleList([torch.nn.Linear(d_in, d) for _ in range(3)])

    def forward(self, image, dna, language):
        f = torch.nn.functional.normalize
        return (f(self.enc[0](image), dim=-1), f(self.enc[1](dna), dim=-1),
                f(self.enc[2](language["input_ids"].float()), dim=-1), torch.tensor(14.3), None)


def _toy_loader(n, d_in, batch, seed, n_species=12):
    gen = torch.Generator().manual_seed(seed)
    out = []
    for b0 in range(0, n, batch):
        m = min(batch, n - b0)
        sp = torch.randint(0, n_species, (m,), generator=gen)
        base = torch.nn.functional.one_hot(sp, d_in).float() * 3
        labels = {"order": [f"o{int(s) % 2}" for s in sp], "family": [f"f{int(s) % 3}" for s in sp],
                  "genus": [f"g{int(s) % 6}" for s in sp], "species": [f"s{int(s)}" for s in sp]}
        ids = (base + 0.3 * torch.randn(m, d_in, generator=gen))
        out.append(([f"id{b0 + i}" for i in range(m)], base + 0.3 * torch.randn(m, d_in, generator=gen),
                    base + 0.3 * torch.randn(m, d_in, generator=gen), ids, torch.zeros(m, d_in), torch.ones(m, d_in),
                    labels))
    return out


def test_device_resident_eval_pipeline_matches_oracle():
    """get_features_and_label (device) -> inference_and_print_result == the numpy restatement of
    inference_epoch.py:42-125 + util.py:702-742 + util.py:601-700 on the same toy model."""
    import clibd_b200 as cb
    from clibd_b200 import seam
    torch.manual_seed(0)
    d_in, d = 16, 32
    model = _ToyModel(d_in, d).to(_dev())
    splits = {"key": _toy_loader(90, d_in, 32, 1), "seen": _toy_loader(50, d_in, 16, 2),
              "unseen": _toy_loader(41, d_in, 16, 3)}
    recorded = {}

    class _Recorder(torch.nn.Module):
        def __init__(self, inner, sink):
            super().__init__()
            self.inner, self.sink = inner, sink

        def forward(self, *a):
            out = self.inner(*a)
            self.sink.append([o.detach().cpu() for o in out[:3]])
            return out

    dev_dicts = {}
    for k, v in splits.items():
        recorded[k] = []
        dev_dicts[k] = seam.get_features_and_label(v, _Recorder(model, recorded[k]), _dev(), for_key_set=(k == "key"))

    def ref_split(loader, for_key_set, recorded):
        # the numpy restatement consumes the very outputs the model produced during the device run (two GEMM
        # launches of the same shape need not agree bit for bit)
        feats = [[], [], []]
        labels, names = [], []
        for (pid, img, dna, ids, tt, am, lab), outs in zip(loader, recorded):
            for lst, o in zip(feats, outs):
                lst.append(so.f_normalize_rows(o.numpy()))
            labels.extend(seam.convert_label_dict_to_list_of_dict(lab))
            names.extend(pid)
        img, dna, txt = (np.concatenate(f) for f in feats)
        dct = {"file_name_list": names, "label_list": labels, "encoded_image_feature": img,
               "encoded_dna_feature": dna, "encoded_language_feature": txt}
        dct.update(ko.derived_feature_types(img, dna, txt, for_key_set=for_key_set, labels=labels))
        return dct

    ref = {k: ref_split(v, k == "key", recorded[k]) for k, v in splits.items()}
    for k in splits:
        assert dev_dicts[k]["label_list"] == ref[k]["label_list"]
        assert dev_dicts[k]["file_name_list"] == ref[k]["file_name_list"]
        host = {}
        for ft in ("encoded_image_feature", "encoded_dna_feature", "encoded_language_feature"):
            host[ft] = dev_dicts[k][ft].cpu().numpy().astype(np.float64)
            np.testing.assert_allclose(host[ft], ref[k][ft], rtol=5e-7, atol=0)  # float32 rounding of the normalise
        # the derived types are exact functions of the stored features (averaging cancels, so they are compared
        # against the restatement applied to the SAME stored values, bit for bit)
        drv = ko.derived_feature_types(host["encoded_image_feature"], host["encoded_dna_feature"],
                                       host["encoded_language_feature"], for_key_set=(k == "key"),
                                       labels=dev_dicts[k]["label_list"])
        for ft in ("averaged_feature", "concatenated_feature", "all_key_features"):
            if drv[ft] is None:
                assert dev_dicts[k][ft] is None
            else:
                assert np.array_equal(dev_dicts[k][ft].cpu().numpy().astype(np.float64), drv[ft]), ft
    assert dev_dicts["key"]["all_key_features"].shape == (270, d)
    assert dev_dicts["key"]["all_key_features_label"] == ref["key"]["all_key_features_label"]
    assert dev_dicts["seen"]["all_key_features"] is None

    k_list = [1, 3, 5]
    acc, per_class, pred = cb.inference_and_print_result(dev_dicts["key"], dev_dicts["seen"], dev_dicts["unseen"],
                                                         args=None, k_list=k_list, verbose=False)
    # the oracle searches the DEVICE pipeline's own features (bit-identical inputs), so indices must agree exactly
    for qt in ("encoded_image_feature", "averaged_feature", "concatenated_feature"):
        for kt in ("encoded_dna_feature", "averaged_feature", "concatenated_feature", "all_key_features"):
            qf = dev_dicts["seen"][qt].cpu().numpy()
            kf = dev_dicts["key"][kt].cpu().numpy()
            if qf.shape[1] != kf.shape[1]:
                assert acc[qt][kt] == {}
                continue
            kl = dev_dicts["key"]["all_key_features_label"] if kt == "all_key_features" else None
            if kl is None:
                # util.py:651-652: once all_key_features was visited its label list sticks for later key types; the
                # key types iterated BEFORE it use the plain label list
                kl = dev_dicts["key"]["label_list"]
            rp = ko.make_prediction(qf, kf, kl, max_k=5)
            assert pred[qt][kt]["curr_seen_pred_list"] == rp
            assert acc[qt][kt]["seen"]["micro_acc"] == ko.micro_accuracy_ref_style(rp, dev_dicts["seen"]["label_list"], k_list)
