"""Mixture assembly (separation/look2hear/datas/movingdatamodule.py:105-124): the oracle against golden
vectors produced by exec-ing the reference's own lines, and (-m gpu) the CUDA path against both."""
import numpy as np
import pytest
import torch

from oracle import sonicsim_oracle as so


def test_oracle_matches_reference_lines_bit_exact(golden):
    g = golden("mix_stems")
    for k in range(int(g["n_cases"])):
        mix, spk = so.mix_stems(torch.from_numpy(g[f"spk{k}"]), torch.from_numpy(g[f"noise{k}"]), g[f"sirs{k}"], g[f"snr{k}"])
        assert np.array_equal(mix.numpy(), g[f"mix{k}"]) and np.array_equal(spk.numpy(), g[f"spk_out{k}"])


def _fo_data(g, k):
    pts = g[f"fo_points{k}"]
    data = {f"spk{s}": {"start_end_points": pts[s].tolist()} for s in range(pts.shape[0])}
    data["noise"] = {"other": 1}
    return data


def test_oracle_enhancement_variant_bit_exact(golden):
    """overlap_audio / find_overlap_region / the noisy-mixture block of the enhancement dataloader
    (enhancement/look2hear/datas/movingdatamodule.py:34-75, 235-257)."""
    import ast
    import random
    g = golden("mix_noisy")
    for k in range(int(g["n_ov"])):
        assert np.array_equal(so.overlap_audio(g[f"ov_x{k}"], int(g[f"ov_sr{k}"]), float(g[f"ov_delay{k}"])), g[f"ov_y{k}"])
    for k in range(int(g["n_fo"])):
        random.seed(int(g[f"fo_seed{k}"]))
        assert so.find_overlap_region(_fo_data(g, k), **ast.literal_eval(str(g[f"fo_kw{k}"]))) == tuple(g[f"fo_out{k}"])
    for k in range(int(g["n_mn"])):
        m = so.mix_noisy(torch.from_numpy(g[f"mn_spk{k}"]), torch.from_numpy(g[f"mn_noise{k}"]), g[f"mn_snr{k}"], int(g[f"mn_sr{k}"]))
        assert np.array_equal(m.numpy(), g[f"mn_mix{k}"])


def test_find_overlap_region_dropin_same_random_stream(golden):
    import ast
    import random
    from sonicsim_b200 import mix as smix
    g = golden("mix_noisy")
    for k in range(int(g["n_fo"])):
        random.seed(int(g[f"fo_seed{k}"]))
        assert smix.find_overlap_region(_fo_data(g, k), **ast.literal_eval(str(g[f"fo_kw{k}"]))) == tuple(g[f"fo_out{k}"])
        after = random.random()
        random.seed(int(g[f"fo_seed{k}"]))
        so.find_overlap_region(_fo_data(g, k), **ast.literal_eval(str(g[f"fo_kw{k}"])))
        assert random.random() == after                        # consumed exactly as many draws
    with pytest.raises(ValueError):
        smix.find_overlap_region({"noise": {}})


@pytest.mark.gpu
def test_gpu_enhancement_variant_matches_golden(golden):
    from sonicsim_b200 import mix as smix
    g = golden("mix_noisy")
    for k in range(int(g["n_ov"])):
        y = smix.overlap_audio(torch.from_numpy(g[f"ov_x{k}"]), int(g[f"ov_sr{k}"]), delay=float(g[f"ov_delay{k}"]))
        assert np.array_equal(y.numpy(), g[f"ov_y{k}"])        # additions in the reference's order: bit-exact
    for k in range(int(g["n_mn"])):
        m = smix.mix_noisy(torch.from_numpy(g[f"mn_spk{k}"]), torch.from_numpy(g[f"mn_noise{k}"]), g[f"mn_snr{k}"],
                           sample_rate=int(g[f"mn_sr{k}"]))
        assert m.shape == torch.Size(g[f"mn_mix{k}"].shape)
        assert so.rel_rms(m.numpy(), g[f"mn_mix{k}"]) < 1e-5
    # seeded draw like :244
    torch.manual_seed(3)
    m1 = smix.mix_noisy(torch.from_numpy(g["mn_spk0"]), torch.from_numpy(g["mn_noise0"]), sample_rate=int(g["mn_sr0"]))
    torch.manual_seed(3)
    snr = torch.Tensor(1).uniform_(-10, 15).numpy()
    m2 = so.mix_noisy(torch.from_numpy(g["mn_spk0"]), torch.from_numpy(g["mn_noise0"]), snr, int(g["mn_sr0"]))
    assert so.rel_rms(m1.numpy(), m2.numpy()) < 1e-5
    # delay longer than the clip: both shifted copies vanish
    x = torch.from_numpy(g["ov_x2"])
    assert np.array_equal(smix.overlap_audio(x, 16000, delay=6).numpy(), g["ov_x2"])


@pytest.mark.gpu
def test_gpu_mix_matches_golden(golden):
    from sonicsim_b200 import mix as smix
    g = golden("mix_stems")
    for k in range(int(g["n_cases"])):
        mix, spk = smix.mix_stems(torch.from_numpy(g[f"spk{k}"]), torch.from_numpy(g[f"noise{k}"]), g[f"sirs{k}"], g[f"snr{k}"])
        assert mix.shape == torch.Size(g[f"mix{k}"].shape) and spk.shape == torch.Size(g[f"spk_out{k}"].shape)
        assert so.rel_rms(mix.numpy(), g[f"mix{k}"]) < 1e-5
        assert so.rel_rms(spk.numpy(), g[f"spk_out{k}"]) < 1e-5


@pytest.mark.gpu
def test_gpu_mix_seeded_rng_and_clamp():
    from sonicsim_b200 import mix as smix
    rng = np.random.default_rng(9)
    spk = torch.from_numpy((rng.standard_normal((3, 2, 64000)) * 0.1).astype(np.float32))
    spk[2] *= 1e-5                                            # gain clamps at +40 dB
    noi = torch.from_numpy((rng.standard_normal((2, 2, 64000)) * 0.02).astype(np.float32))
    torch.manual_seed(4)
    m1, s1 = smix.mix_stems(spk, noi)
    torch.manual_seed(4)
    sirs = torch.Tensor(2).uniform_(-6, 6).numpy()
    snr = torch.Tensor(1).uniform_(10, 20).numpy()
    m2, s2 = so.mix_stems(spk, noi, sirs, snr)
    assert so.rel_rms(m1.numpy(), m2.numpy()) < 1e-5 and so.rel_rms(s1.numpy(), s2.numpy()) < 1e-5
    assert abs(float(s1[2].abs().max() / spk[2].abs().max()) - 100.0) < 1e-3
    # silence everywhere: max(1e-20, .) floor keeps everything finite
    z = torch.zeros(2, 8000)
    m, s = smix.mix_stems(z, torch.zeros(1, 8000), [0.0], 15.0)
    assert torch.isfinite(m).all() and not m.any()


@pytest.mark.gpu
def test_gpu_render_mixtures_whole_path_on_device():
    """render (moving speakers + static noises) -> SIR/SNR mixture, all in HBM, vs oracle render + oracle mix."""
    from sonicsim_b200 import render
    rng = np.random.default_rng(31)
    N, C, L, P = 48000, 2, 1200, 5
    utts = []
    for _ in range(2):
        utts.append({"speakers": [(so.synth_dry(rng, N), so.synth_rirs(rng, P, C, L), so.synth_path(rng, P)) for _ in range(2)],
                     "noises": [(so.synth_dry(rng, N) * 0.3, so.synth_rirs(rng, 1, C, L)[0])],
                     "sirs": np.array([rng.uniform(-6, 6)], np.float32), "snr": float(rng.uniform(10, 20))})
    np.random.seed(77)
    got = render.render_mixtures(utts)
    np.random.seed(77)
    for u, g in zip(utts, got):
        stems = []
        for x, h, pos in u["speakers"]:
            idx, w = so.setup_dynamic_interp(pos, N)
            stems.append(so.convolve_moving_receiver(x, h, idx, w))
        noises = [so.convolve_fixed_receiver(x[None], h) for x, h in u["noises"]]
        mix, spk = so.mix_stems(torch.from_numpy(np.stack(stems)), torch.from_numpy(np.stack(noises)), u["sirs"], u["snr"])
        assert g["mix"].shape == (C, N)
        assert so.rel_rms(g["mix"], mix.numpy()) < 1e-4
        assert so.rel_rms(g["speakers"], spk.numpy()) < 1e-4
