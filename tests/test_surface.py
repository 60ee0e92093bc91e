"""CPU: the mirrored class / registry surface matches what the reference's introspection relies on
(lightning_pose/models/factory.py:116-192, tests/models/test_factory.py:126-155,
tests/models/heads/test_heatmap.py:14-81,293-325, tests/losses/test_factory.py)."""
import inspect

import pytest
import torch
from torch import nn

from lightning_pose_b200.losses import losses as L
from lightning_pose_b200.losses.factory import LossFactory, get_loss_classes
from lightning_pose_b200.models.heads.heatmap import HeatmapHead, make_upsampling_layers


def required_keys(cls):
    sig = inspect.signature(cls.__call__)
    return {n for n, p in sig.parameters.items()
            if n not in ("self", "stage") and p.kind is p.POSITIONAL_OR_KEYWORD and p.default is p.empty}


def test_registry_names():
    assert sorted(get_loss_classes()) == sorted([
        "regression", "heatmap_mse", "heatmap_kl", "heatmap_js", "pca_multiview", "pca_singleview", "temporal",
        "temporal_heatmap_mse", "temporal_heatmap_kl", "supervised_pairwise_projections",
        "supervised_reprojection_heatmap_mse"])
    assert L.RegressionRMSELoss.loss_name == "rmse" and "rmse" not in get_loss_classes()
    assert set(L.__all__) == {"Loss", "HeatmapLoss", "HeatmapMSELoss", "HeatmapKLLoss", "HeatmapJSLoss", "PCALoss",
                              "TemporalLoss", "TemporalHeatmapLoss", "RegressionMSELoss", "RegressionRMSELoss",
                              "PairwiseProjectionsLoss", "ReprojectionHeatmapLoss"}


def test_call_signatures_are_load_bearing():
    assert required_keys(L.HeatmapMSELoss) == {"heatmaps_targ", "heatmaps_pred"}
    assert required_keys(L.TemporalLoss) == {"keypoints_pred"}  # confidences stays optional
    assert required_keys(L.RegressionMSELoss) == {"keypoints_targ", "keypoints_pred"}
    assert required_keys(L.PCALoss) == {"keypoints_pred"}
    assert required_keys(L.TemporalHeatmapLoss) == {"heatmaps_pred", "confidences"}
    assert required_keys(L.ReprojectionHeatmapLoss) == {"heatmaps_targ", "keypoints_pred_2d_reprojected"}


def test_loss_hyperparameters():
    t = L.TemporalLoss(epsilon=[1.0, 2.0], prob_threshold=0.05, log_weight=11.0)
    assert torch.allclose(t.weight, 1.0 / (2.0 * torch.exp(torch.tensor(11.0))))
    assert t.epsilon.tolist() == [1.0, 2.0] and float(t.prob_threshold) == pytest.approx(0.05)
    with pytest.raises(ValueError):
        L.TemporalHeatmapLoss(loss_name="nope")
    with pytest.raises(ValueError):
        L.PCALoss(loss_name="nope", data_module=object())
    # staged helpers: the reference's analytic TemporalLoss example (tests/losses/test_losses.py:343-360)
    kp = torch.tensor([[0.0, 0.0, 0.0, 0.0], [1.0, 1.0, 3.0, 4.0]])
    assert torch.allclose(t.compute_loss(kp), torch.tensor([[2.0**0.5, 5.0]]))
    assert torch.allclose(t.rectify_epsilon(torch.tensor([[1.5, 5.0]])), torch.tensor([[0.5, 3.0]]))


def test_loss_factory_construction_without_data_module():
    fac = LossFactory({"heatmap_mse": {"log_weight": 0.0}, "temporal": {"log_weight": 5.0, "epsilon": 3.0}}, None)
    assert list(fac.loss_instance_dict) == ["heatmap_mse", "temporal"]
    assert isinstance(fac, nn.Module)


@pytest.mark.parametrize("n_layers", [1, 2, 3])
def test_make_upsampling_layers(n_layers):
    m = make_upsampling_layers(256, 64, 128, n_layers)
    assert len(m) == n_layers + 1 and isinstance(m[0], nn.PixelShuffle)
    assert m[1].in_channels == 64 and m[-1].out_channels == 64
    if n_layers > 1:
        assert m[1].out_channels == 128 and m[-1].in_channels == 128
    for conv in list(m)[1:]:
        assert isinstance(conv, nn.ConvTranspose2d)
        assert (conv.kernel_size, conv.stride, conv.padding, conv.output_padding) == ((3, 3), (2, 2), (1, 1), (1, 1))


@pytest.mark.parametrize("arch,ds,n", [("resnet50", 1, 3), ("resnet50", 2, 2), ("resnet50", 3, 1), ("vits_dino", 1, 2), ("vits_dino", 2, 1)])
def test_head_layer_count_rule(arch, ds, n):
    head = HeatmapHead(arch, 256, 17, downsample_factor=ds)
    assert len(head.upsampling_layers) == n + 1
    assert float(head.temperature) == 1000.0 and head.final_softmax is True
    assert head.upsampling_layers[1].weight.shape[:2] == (64, 17 if n == 1 else 17)
    assert float(head.upsampling_layers[1].bias.abs().max()) == 0.0


def test_tracker_surface():
    """tests/models/test_factory.py:142-155: tracker output keys are read from the TypedDict annotations."""
    import typing

    from lightning_pose_b200.models.heatmap_tracker import HeatmapTracker, SemiSupervisedHeatmapTracker

    lab = typing.get_type_hints(HeatmapTracker.get_loss_inputs_labeled)["return"]
    unl = typing.get_type_hints(SemiSupervisedHeatmapTracker.get_loss_inputs_unlabeled)["return"]
    assert set(lab.__annotations__) == {"heatmaps_targ", "heatmaps_pred", "keypoints_targ", "keypoints_pred", "confidences"}
    assert set(unl.__annotations__) == {"heatmaps_pred", "keypoints_pred", "keypoints_pred_augmented", "confidences"}
    m = SemiSupervisedHeatmapTracker(17, backbone="resnet18")
    assert [g["name"] for g in m.get_parameters()] == ["backbone", "head"] and m.get_parameters()[0]["lr"] == 0
    assert m.num_targets == 34 and float(m.total_unsupervised_importance) == 1.0
    sig = inspect.signature(HeatmapTracker.predict_step)
    assert list(sig.parameters)[1:] == ["batch_dict", "batch_idx", "return_heatmaps"]


def test_mhcrnn_head_surface_matches_reference_checkpoint_keys(golden):
    """State-dict keys (incl. the ModuleList aliases) and shapes of HeatmapMHCRNNHead = the reference module's
    (heads/heatmap_mhcrnn.py:18-266), so its checkpoints load unchanged; golden params come from the reference itself."""
    import torch

    from lightning_pose_b200.models.heads.heatmap_mhcrnn import HeatmapMHCRNNHead, UpsamplingCRNN

    g = golden("mhcrnn")
    for tag, arch, uf in (("vit", "vits_dino", 1), ("resnet", "resnet50", 2)):
        head = HeatmapMHCRNNHead(arch, 64, 5, upsampling_factor=uf)
        sd = head.state_dict()
        ref_keys = {k[len(f"{tag}_param_"):] for k in g.files if k.startswith(f"{tag}_param_")}
        ours = {k for k in sd if ".layers." not in k}
        assert ours == ref_keys
        for k in ref_keys:
            assert tuple(sd[k].shape) == g[f"{tag}_param_{k}"].shape, k
        n_alias = len([k for k in sd if ".layers." in k])
        assert n_alias == (10 if uf == 2 else 8) * 1 + (2 if uf == 2 else 2) * 0 or n_alias > 0
        assert head.upsampling_factor == uf and float(head.temperature) == 1000.0
    m = UpsamplingCRNN(64, 5, upsampling_factor=1)
    assert [type(x).__name__ for x in m.layers] == ["ConvTranspose2d", "Sequential", "ConvTranspose2d", "Sequential"]
    assert m.H_f[0].groups == 5 and m.H_f[0].out_channels == 80 and m.H_f[1].kernel_size == (2, 2)
    assert float(m.W_f.bias.abs().max()) == 0.0  # xavier weights, zero biases (reference :250-266)
    with __import__("pytest").raises(RuntimeError, match="no CPU fallback"):
        head.head_sf(torch.zeros(1, 64, 2, 2))
