"""Host-side half of the drop-in proof (no GPU): with `sparf_b200.install_as_reference_renderer()` the reference's own
trainer module resolves `source.models.renderer.Graph` to this package's mirror, i.e. `joint_pose_nerf_trainer.Graph`
(the subclass the reference's trainer instantiates, joint_pose_nerf_trainer.py:710) derives from OUR `Graph`.  The
GPU half (the subclass + the reference's loss modules actually running on the kernels, against the goldens) is
tests/test_losses.py::test_reference_trainer_graph_and_losses_on_our_renderer."""
import sys

import pytest


def test_reference_trainer_graph_subclasses_our_graph():
    from oracle import ref_loader
    if not ref_loader.ref_root():
        pytest.skip("reference not available (neither /root/reference nor oracle/_ref)")
    import sparf_b200.frequency_nerf
    import sparf_b200.renderer
    try:
        mods = ref_loader.load(stack="trainer", shadow_renderer=True)
        assert sys.modules["source.models.renderer"] is sparf_b200.renderer
        assert sys.modules["source.models.frequency_nerf"] is sparf_b200.frequency_nerf
        assert mods.renderer is sparf_b200.renderer
        assert sparf_b200.renderer.Graph in mods.joint.Graph.__mro__
        # the reference's loss factory is the reference's own code
        assert mods.loss_factory.__file__.startswith(mods.root)
        # and without the swap the same loader gives the reference's own renderer back (no leak between the two)
        mods2 = ref_loader.load(stack="trainer", shadow_renderer=False)
        assert mods2.renderer is not sparf_b200.renderer
        assert sparf_b200.renderer.Graph not in mods2.joint.Graph.__mro__
    finally:
        ref_loader._purge()
