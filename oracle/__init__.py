"""CPU oracle: a restatement of google/compare_gan's GAN-step / FID hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``compare_gan_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and there only as the checker
(or as the timed CPU arm), never as the thing shipped.

The reference (TF 1.x graph code) cannot be imported in this image (no
tensorflow / gin / tfgan / tfhub, Python 3.12), so this is a PyTorch-CPU fp32 +
NumPy fp64 restatement that follows the reference file by file; every function
cites the reference ``file:line`` it restates (paths relative to
``/root/reference/compare_gan/``).

Parity pinning (tests/test_oracle_golden.py):
  * BatchNorm 4x2x1x3 golden tensor        architectures/arch_ops_test.py:32-61
  * BN accumulator state machine           architectures/arch_ops_test.py:63-132
  * cross-replica mean vectors             tpu/tpu_ops_test.py:79-83
  * 2-replica sync-BN == 1-replica golden  architectures/arch_ops_tpu_test.py:112-133
  * FID = 89.091                           metrics/fid_score_test.py:31-40
  * resnet_cifar variable names/shapes     architectures/resnet_norm_test.py
  * BigGAN128 parameter counts             architectures/resnet_biggan_test.py:139,154
  * D/G step-counter rule                  gans/modular_gan_test.py:175-177
PARITY UNPINNED (no golden in the reference; TF semantics restated from the
call sites + App. A of SURVEY.md, cross-checked by finite differences):
conv2d / deconv2d / spectral_norm / non_local_block numerics, the losses, the
penalties, Inception Score, KID, and the Inception network weights (not
vendored, not downloadable).
"""
