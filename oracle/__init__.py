"""CPU oracle for the video-dqn Q-learning hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product path.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and there only as the checker / the reported CPU
baseline -- never as the thing measured as "ours" or shipped.

Pinning status: the reference repository has no tests, golden
vectors or known-answer fixtures, and every FLOP of the path is executed by
third-party libraries (torch==1.3.1, torchvision==0.4.2, pinned in the
reference's requirements.txt:138,140).  The oracle is therefore pinned against
OUTPUTS OF THE REFERENCE ITSELF run in the build container
(``oracle/make_goldens.py`` imports ``/root/reference/archs/HabitatDQNMultiAction.py``
and runs the reference's own ``run_train`` for three steps) and the resulting
vectors are committed under ``tests/golden/``.

Files: ``qstep.py`` (Q-learning step, both architectures), ``inverse.py`` (inverse-dynamics model: labelling
forward and training step), ``td_adam_ref.c`` (plain-C restatement of the TD loss and Adam, built by
``oracle/Makefile`` into ``oracle/_ref/libtdref.so``: a torch-free second oracle pinned to the same vectors),
``make_*goldens.py`` (the generators, each executing the reference's own code), ``crosscheck_run_train.py``
(the reference's own ``run_train`` loop run live against the oracle), ``probe_basic_bf16.py``
(what PyTorch's own bf16 autocast does to the train-mode-BatchNorm step: the yardstick for that path's bars).
"""
