"""ONE handle, several GPUs (the reference's caller holds one handle: yoloface.c:216-240).
yf_b200_config.device_mask / YF_B200_DEVICES make ai_network_run / yf_b200_run / yf_b200_detect split the images of a
call into contiguous ranges, one per GPU (worker thread + streams per GPU inside the library, no collective); results
land in the caller's buffers at the ranges' offsets.  Needs two GPUs: run with `gpurun --gpus 2` (the 1-GPU box skips)."""
import os

import numpy as np
import pytest

import pkg

pytestmark = pytest.mark.gpu


def ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:  # noqa: BLE001
        return 0


@pytest.fixture(scope="module")
def yf():
    return pkg.load()


def batch(golden, n, seed):
    rng = np.random.default_rng(seed)
    x = rng.integers(-128, 128, (n, 56, 56, 3), dtype=np.int8)
    x[::2] = golden["images"][np.arange(len(x[::2])) % 27]
    return x


@pytest.mark.skipif(ngpu() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_one_handle_many_gpus_bit_exact(yf, oracle, golden):
    devs = list(range(min(ngpu(), 8)))
    net = yf.Network(devices=devs, chunk_images=512)
    one = yf.Network(device=0, chunk_images=512)
    try:
        for n, seed in ((2 * len(devs), 1), (1000, 2), (4099, 3)):       # ragged splits included
            x = batch(golden, n, seed)
            want = oracle.run_batch(x, threads=os.cpu_count())
            assert np.array_equal(net.run(x), want)                       # yf_b200_run
            if n <= 65535:
                assert np.array_equal(net.ai_run(x), want)                # ai_network_run, n_batches = n
            dets, counts = net.detect(x, 0.7, 0.4, max_det=16)            # decode + NMS on every device
            d1, c1 = one.detect(x, 0.7, 0.4, max_det=16)
            assert np.array_equal(counts, c1) and np.array_equal(dets, d1)
        st = net.stats()
        assert st["images"] >= 2 * (2 * len(devs) + 1000 + 4099)
        # a call too small to split, and a device pointer, run on one member
        x = batch(golden, 1, 9)
        assert np.array_equal(net.run(x), oracle.run_batch(x, threads=1))
        import torch
        xd = torch.from_numpy(batch(golden, 64, 10)).cuda(devs[-1])
        yd = torch.empty((64, 7, 7, 18), dtype=torch.int8, device=xd.device)
        net.run(xd, yd, n=64)
        torch.cuda.synchronize(xd.device)
        assert np.array_equal(yd.cpu().numpy(), oracle.run_batch(xd.cpu().numpy(), threads=4))
        # settings reach every member
        net.set_input_size(112, 112)
        x2 = np.random.default_rng(5).integers(-128, 128, (40, 112, 112, 3), dtype=np.int8)
        out = net.run(x2)
        for i in (0, 13, 39):
            assert np.array_equal(out[i], oracle.run(x2[i]))
        assert net.get_error() == (0, 0)
    finally:
        net.close(); one.close()


@pytest.mark.skipif(ngpu() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_env_selects_all_devices(yf, oracle, golden, monkeypatch):
    monkeypatch.setenv("YF_B200_DEVICES", "all")
    net = yf.Network(chunk_images=256)
    try:
        x = batch(golden, 777, 4)
        assert np.array_equal(net.run(x), oracle.run_batch(x, threads=os.cpu_count()))
    finally:
        net.close()


def test_single_gpu_mask_is_a_plain_context(yf, oracle, golden):
    net = yf.Network(devices=[0], chunk_images=128)
    try:
        x = batch(golden, 130, 6)
        assert np.array_equal(net.run(x), oracle.run_batch(x, threads=os.cpu_count()))
    finally:
        net.close()
