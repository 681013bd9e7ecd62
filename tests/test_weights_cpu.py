"""tch `.model` import / export (net6.rs:87-96): the creation order is recovered from tch's
`name__<count>` suffixes whatever order the file was written in and whichever order a conv / BatchNorm registers its
variables in.  (No real `.model` ships with the reference: parity-unpinned against tch itself, see weights.py.)"""
import numpy as np
import pytest

from tak_b200 import weights as W


@pytest.mark.parametrize("arch", [5, 6])
@pytest.mark.parametrize("conv_bias_first,bn_stats_first", [(True, True), (True, False), (False, True), (False, False)])
def test_model_file_round_trip(tmp_path, arch, conv_bias_first, bn_stats_first):
    blob = W.random_weights(arch, seed=3)
    path = str(tmp_path / "net.model")
    W.save_tch_model(blob, arch, path, conv_bias_first, bn_stats_first, shuffle_seed=arch)
    back = W.load_tch_model(path, arch)
    assert back.dtype == np.float32 and np.array_equal(back, blob)


def test_variable_names_follow_the_suffix_rule():
    names = [n for n, _ in W.tch_variable_names(6)]
    # conv2d registers bias then weight, batch_norm2d its running statistics then gamma / beta (tch 0.7.2)
    assert names[:8] == ["bias", "weight", "running_mean", "running_var", "weight__4", "bias__5", "bias__6", "weight__7"]
    assert len(names) == len(set(names)) == len(W.spec(6)) == 2 + 4 + 16 * 12 + 4
    order = W._creation_order(list(reversed(names)))
    assert [order[n] for n in names] == list(range(len(names)))


def test_wrong_architecture_is_rejected(tmp_path):
    path = str(tmp_path / "net5.model")
    W.save_tch_model(W.random_weights(5, seed=1), 5, path)
    with pytest.raises(ValueError):
        W.load_tch_model(path, 6)
    with pytest.raises(ValueError):
        W._creation_order(["weight", "bias", "gamma"])
