"""CPU oracle for the SGAM per-frame hot path.  TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a CPU restatement of the reference's algorithm
(yshen47/SGAM_NeurIPS22 @ 780feff) used as the *checker* for the CUDA path.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.  The product
package ``sgam_neurips22_b200`` and the drop-in ``sgam`` package never do.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the unmodified
reference from ``/root/reference`` (in the build container), runs it on seeded
inputs and commits the outputs under ``tests/golden/``; ``tests/test_oracle_golden.py``
checks every oracle function against those vectors (bit-exact for the integer /
index / mask work and the splat arithmetic, 1e-5 for the conv network which
re-uses the same torch CPU fp32 operators as the reference).

Layout
  csrc/oracle.c   plain C restatement of the byte/integer/index work
                  (splat, median fill, hole mask, inverse warp + z-test,
                  VQ nearest neighbour, depth coding, uint8 packing)
  native.py       ctypes bindings + build recipe for csrc/oracle.c
  network.py      torch fp32 functional restatement of the VQGAN
                  Encoder / Decoder / ResnetBlock / AttnBlock
  model.py        VQModel.get_x / encode / decode / forward restatement
  recipes.py      seeded synthetic inputs and weights shared by the golden
                  generator, the oracle tests and the GPU parity tests
"""
