"""CPU: host-side policy of the Mapper drop-in (naruto_b200/coslam_mapper.py) with the GPU objects replaced by stand-ins --
when a batch size gets its CUDA graphs, how many MappingStep objects are kept."""
import types

from naruto_b200 import coslam_mapper as cm


class _FakeStep:
    made = 0

    def __init__(self, plan, cfg, n_rays, dev, process_group=None, use_graph=True, state=None):
        _FakeStep.made += 1
        self.B, self.use_graph, self.released = n_rays, use_graph, False

    def release_graphs(self):
        self.released = True


def _mapper(monkeypatch, use_graph=True):
    monkeypatch.setattr(cm, 'MappingStep', _FakeStep)
    fm = object.__new__(cm.FusedMapper)
    fm.plan, fm.cfg, fm.dev, fm.pg, fm.use_graph, fm.state = None, {}, 'cuda', None, use_graph, None
    fm.steps, fm.uses, fm._seed = {}, {}, 0
    return fm


def test_a_new_batch_size_runs_eagerly_and_is_captured_when_it_comes_back(monkeypatch):
    fm = _mapper(monkeypatch)
    a = fm.step_for(2148)
    assert a.use_graph is False, 'first global_BA call of a size: plain launches'
    assert fm.step_for(2148) is a and a.use_graph is True, 'second call of the same size: graphs'
    b = fm.step_for(2560)                       # the key-frame database grew: another size
    assert b is not a and b.use_graph is False
    # first_frame_mapping announces its 200 iterations: captured at once
    c = fm.step_for(2048, n_iters=200)
    assert c.use_graph is True
    # a caller that asked for no graphs never gets them
    fm2 = _mapper(monkeypatch, use_graph=False)
    d = fm2.step_for(64)
    fm2.step_for(64)
    assert d.use_graph is False


def test_at_most_eight_batch_sizes_are_kept(monkeypatch):
    fm = _mapper(monkeypatch)
    steps = [fm.step_for(1000 + i) for i in range(10)]
    assert len(fm.steps) == 8 and steps[0].released and steps[1].released and not steps[2].released
    assert 1000 not in fm.uses and 1009 in fm.uses
    fm.release()
    assert not fm.steps and all(s.released for s in steps[2:])
    assert fm.next_seed() == 1 and fm.next_seed() == 2
