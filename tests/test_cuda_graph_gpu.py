"""The boundary contract says the library neither allocates nor synchronises after ub_plan_create
and launches only on the caller's stream (its internal weight-gradient stream is forked from and
joined back into it with events), so a forward pass — and a whole training step — can be captured
into a CUDA graph and replayed. Replays must reproduce eager execution
bit for bit (all reductions are in fixed order)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import unet_ref  # noqa: E402


def _model(seed, train):
    from unet_segmentation_b200.unet import UNet

    sd = unet_ref.make_state_dict(1, 2, seed=seed)
    gen = torch.Generator().manual_seed(17)
    for k in [k for k in sd if k.endswith("running_mean")]:
        nf = sd[k].numel()
        sd[k] = torch.randn(nf, generator=gen) * 0.1
        sd[k.replace("running_mean", "running_var")] = 0.5 + torch.rand(nf, generator=gen)
    m = UNet(1, 2)
    m.load_state_dict(sd)
    m = m.cuda()
    return m.train() if train else m.eval()


def _warm_up(fn, times=2):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(times):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()


def test_eval_forward_replays_from_a_cuda_graph():
    model = _model(3, train=False)
    x_static, _, _ = unet_ref.synthetic_batch(1, size=316, seed=1, device="cuda")
    with torch.no_grad():
        _warm_up(lambda: model.predict_mask(x_static))
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            logits_s, mask_s = model.predict_mask(x_static)
        for seed in (2, 3):
            new, _, _ = unet_ref.synthetic_batch(1, size=316, seed=seed, device="cuda")
            x_static.copy_(new)
            graph.replay()
            torch.cuda.synchronize()
            got_logits, got_mask = logits_s.clone(), mask_s.clone()
            eager_logits, eager_mask = model.predict_mask(new)
            torch.cuda.synchronize()
            assert torch.equal(got_logits, eager_logits) and torch.equal(got_mask, eager_mask)
    assert float(got_logits.abs().max()) > 0


def test_training_step_replays_from_a_cuda_graph():
    """zero_grad -> forward -> weighted CE -> backward (incl. the side-stream weight gradients) ->
    FusedSGD step captured once, replayed twice; losses equal an eager twin's, step by step."""
    from unet_segmentation_b200.loss import WeightedCrossEntropyLoss
    from unet_segmentation_b200.optim import FusedSGD

    img, t, w = unet_ref.synthetic_batch(1, size=252, seed=9, device="cuda")
    crit = WeightedCrossEntropyLoss()

    def make():
        m = _model(5, train=True)
        return m, FusedSGD(m, lr=1e-3, momentum=0.9)

    def step(m, opt):
        opt.zero_grad(set_to_none=True)
        loss = crit(m(img), t, w)
        loss.backward()
        opt.step()
        return loss

    twin, twin_opt = make()
    eager = [float(step(twin, twin_opt)) for _ in range(3)]

    model, opt = make()
    first = []
    _warm_up(lambda: first.append(float(step(model, opt))), times=1)      # step 1, eager
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        loss_s = step(model, opt)                                          # recorded, not executed
    replayed = []
    for _ in range(2):                                                     # steps 2 and 3
        graph.replay()
        torch.cuda.synchronize()
        replayed.append(float(loss_s))
    assert first[0] == eager[0]
    assert replayed == eager[1:], (replayed, eager)
    assert int(model.inc.double_conv[1].num_batches_tracked) == 3
    for (k, a), (_, b) in zip(model.state_dict().items(), twin.state_dict().items()):
        assert torch.equal(a, b), k
