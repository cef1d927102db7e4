"""GraphedStep: the training iteration replayed as one CUDA graph must be the same arithmetic as the iteration
enqueued kernel by kernel (main.py:170-207), must not advance training while capturing, and must fall back to the
eager step for ragged batches (main.py:156-158 filters broken samples)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _make(seed=0):
    from zeroshotvideoclassification_b200 import video_models as vm
    torch.manual_seed(seed)
    model = vm.get_network(vm.default_opt("r2plus1d_18")).cuda().train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, capturable=True)
    crit = torch.nn.MSELoss()

    def step(X, Z):
        opt.zero_grad(set_to_none=True)
        emb, _ = model(X)
        loss = crit(emb, Z)
        loss.backward()
        opt.step()
        return loss

    return model, opt, step


def _data(B, seed=1):
    g = torch.Generator().manual_seed(seed)
    xs = [torch.randn(B, 1, 3, 8, 64, 64, generator=g) for _ in range(3)]
    cls = F.normalize(torch.randn(51, 300, generator=g))
    zs = [cls[torch.randint(0, 51, (B,), generator=g)] for _ in range(3)]
    return xs, zs


def test_graphed_step_matches_eager_and_keeps_state():
    from zeroshotvideoclassification_b200.graph import GraphedStep
    xs, zs = _data(2)

    model_e, _, step_e = _make()
    eager_losses = [float(step_e(x.cuda(), z.cuda())) for x, z in zip(xs, zs)]

    model_g, opt_g, step_g = _make()
    before = {k: v.detach().clone() for k, v in model_g.state_dict().items()}
    gstep = GraphedStep(step_g, (xs[0], zs[0]), model=model_g, optimizer=opt_g)
    # capturing (warm-up iterations included) left parameters, BN running statistics and Adam state untouched
    for k, v in model_g.state_dict().items():
        assert torch.equal(v, before[k]), k
    for st in opt_g.state.values():
        assert float(st["step"]) == 0.0 and float(st["exp_avg"].abs().max()) == 0.0
    assert gstep.launches_per_replay > 300

    graph_losses = []
    for x, z in zip(xs, zs):
        loss = gstep(x.pin_memory(), z.pin_memory())      # host batches: H2D into the captured buffers
        graph_losses.append(float(loss))
    # deterministic kernels (no atomics): the replay is bit-identical to the eager iteration
    assert graph_losses == eager_losses
    sd_e, sd_g = model_e.state_dict(), model_g.state_dict()
    for k in sd_e:
        assert torch.equal(sd_e[k], sd_g[k]), k
    assert gstep.replays == 3


def test_graphed_step_ragged_batch_runs_eagerly():
    from zeroshotvideoclassification_b200.graph import GraphedStep
    xs, zs = _data(3)
    model, opt, step = _make()
    gstep = GraphedStep(step, (xs[0], zs[0]), model=model, optimizer=opt)
    l_full = float(gstep(xs[0], zs[0]))
    l_ragged = float(gstep(xs[1][:2], zs[1][:2]))        # one sample filtered out
    assert gstep.replays == 1
    assert l_full > 0 and l_ragged > 0 and l_ragged == l_ragged


def test_prefetch_pipeline_matches_direct_calls():
    """prefetch() + step() (next batch uploaded on a copy stream during the current replay) gives the same training
    trajectory as passing each batch to the step directly."""
    from zeroshotvideoclassification_b200.graph import GraphedStep
    xs, zs = _data(2)
    model_a, opt_a, step_a = _make()
    ga = GraphedStep(step_a, (xs[0], zs[0]), model=model_a, optimizer=opt_a)
    direct = [float(ga(x, z)) for x, z in zip(xs, zs)]

    model_b, opt_b, step_b = _make()
    gb = GraphedStep(step_b, (xs[0], zs[0]), model=model_b, optimizer=opt_b)
    pinned = [(x.pin_memory(), z.pin_memory()) for x, z in zip(xs, zs)]
    gb.prefetch(*pinned[0])
    piped = []
    for i in range(3):
        loss = gb()
        if i + 1 < 3:
            gb.prefetch(*pinned[i + 1])
        piped.append(float(loss))
    assert piped == direct


def test_eval_after_graph_replays_sees_the_new_weights():
    """evaluate() (main.py:224-257) after more training through graph replays must use the CURRENT weights and running
    statistics: replays write them through raw pointers, which tensor version counters do not see, so the folded
    inference weights are dropped on every replay (engine.note_weights_changed)."""
    from zeroshotvideoclassification_b200.graph import GraphedStep
    xs, zs = _data(2)
    model, opt, step = _make()
    gstep = GraphedStep(step, (xs[0], zs[0]), model=model, optimizer=opt)

    def evaluate():
        model.eval()
        with torch.no_grad():
            emb, _ = model(xs[2].cuda())
        model.train()
        return emb.clone()

    gstep(xs[0], zs[0])
    e1 = evaluate()
    assert torch.equal(e1, evaluate())                    # cached folded weights: same answer
    for x, z in zip(xs, zs):
        gstep(x, z)                                       # replays only: no Python-side in-place op on any parameter
    e2 = evaluate()
    assert not torch.equal(e1, e2), "eval served stale folded weights after graph replays"
    # and it is the answer a freshly folded model gives
    from zeroshotvideoclassification_b200 import engine
    engine.note_weights_changed()
    assert torch.equal(e2, evaluate())


def test_second_backward_through_the_backbone_raises():
    model, _, _ = _make()
    xs, zs = _data(2)
    emb, _ = model(xs[0].cuda())
    loss = F.mse_loss(emb, zs[0].cuda())
    loss.backward(retain_graph=True)
    with pytest.raises(RuntimeError, match="second time"):
        loss.backward()


def test_graphed_fused_adam_follows_the_lr_schedule():
    """GraphedStep + FusedAdam(model=...): the captured update reads the learning rate from a device scalar that is
    refreshed before every replay, and the captured forward reads the weight images the captured update writes; the
    trajectory equals the eager one with the same optimizer and schedule (main.py:133,374: MultiStepLR)."""
    from zeroshotvideoclassification_b200 import video_models as vm
    from zeroshotvideoclassification_b200.graph import GraphedStep
    from zeroshotvideoclassification_b200.optim import FusedAdam
    xs, zs = _data(2)
    runs = []
    for graphed in (False, True):
        torch.manual_seed(0)
        model = vm.get_network(vm.default_opt("r2plus1d_18")).cuda().train()
        opt = FusedAdam(model.parameters(), lr=1e-2, model=model)
        sched = torch.optim.lr_scheduler.MultiStepLR(opt, [2], gamma=0.01)

        def step(X, Z):
            opt.zero_grad(set_to_none=True)
            emb, _ = model(X)
            loss = F.mse_loss(emb, Z)
            loss.backward()
            opt.step()
            return loss

        fn = GraphedStep(step, (xs[0], zs[0]), model=model, optimizer=opt) if graphed else step
        losses = []
        for it in range(5):
            x, z = xs[it % 3], zs[it % 3]
            losses.append(float(fn(x.cuda(), z.cuda())))
            sched.step()
        runs.append((losses, {k: v.detach().clone() for k, v in model.state_dict().items()}))
    assert runs[0][0] == runs[1][0], (runs[0][0], runs[1][0])
    for k in runs[0][1]:
        assert torch.equal(runs[0][1][k], runs[1][1][k]), k
