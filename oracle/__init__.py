"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the streaming RNN-T inference path.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import it, and only as the checker or the timed CPU
baseline -- never as a fallback for the CUDA path (``libreasr_b200`` raises
when its CUDA library is missing).

Contents
--------
``ref_shim``     stub modules that let the UNMODIFIED reference
                 (``/root/reference/libreasr/lib/models.py``) import in the
                 authoring container (fastai2 / IPython / matplotlib absent).
                 Only usable where ``/root/reference`` exists.
``weights``      deterministic, name-keyed synthetic weight + audio generator
                 (the reference ships no pretrained weights and no fixtures).
``rnnt_oracle``  torch-fp32 CPU restatement of the reference algorithm, every
                 function citing the reference file:line it follows.  Travels
                 to the GPU box (the reference tree does not).
``make_golden``  runs the real reference through the shim and writes
                 ``tests/golden/*.npz`` (the fixtures that pin the oracle).

Parity status: the greedy path is PINNED -- ``tests/test_oracle_golden.py``
checks ``rnnt_oracle`` against fixtures produced by the imported reference
modules (``Transducer.decode_greedy`` / ``transcribe_stream`` /
``Encoder`` / ``Predictor`` / ``Joint``).  Beam search is parity-UNPINNED: the
reference has no beam search.
"""
