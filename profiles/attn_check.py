"""GPU check + timing of the fused attention kernels (csrc/attn_tc.cu) without pytest: per-output errors against float64 on
the TF32-rounded operands, then timings at BigGAN-128's shapes against the composed (bmm -> softmax -> bmm) path.
usage: python profiles/attn_check.py [fwd|bwd|time]"""
import sys
import numpy as np
import torch

sys.path.insert(0, ".")
from compare_gan_b200 import kernels as K, tape  # noqa: E402


def rna(a):
  a = np.ascontiguousarray(a, np.float32)
  return ((a.view(np.uint32) + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def rel(a, b):
  return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-30))


def check(stage, bsz, lq, lk, dk, dv):
  rng = np.random.RandomState(lq + dk)
  th, ph, g, gy = [rna(s * rng.randn(*shape)) for s, shape in ((0.5, (bsz, lq, dk)), (0.5, (bsz, lk, dk)), (1, (bsz, lk, dv)), (1, (bsz, lq, dv)))]
  tt, pt, gt = [torch.from_numpy(a).double().requires_grad_(True) for a in (th, ph, g)]
  ref = torch.bmm(torch.softmax(torch.bmm(tt, pt.transpose(1, 2)), -1), gt)
  ref.backward(torch.from_numpy(gy).double())
  td, pd, gd = [K.from_numpy(a, req=True) for a in (th, ph, g)]
  for t in (td, pd, gd):
    t.tf32 = True
  y = K.attention(td, pd, gd)
  torch.cuda.synchronize()
  out = y.cpu()
  print("%-4s B=%d lq=%d lk=%d dk=%d dv=%d: out err %.2e  (nan %d)" % (stage, bsz, lq, lk, dk, dv, rel(out, ref.detach().numpy()), int(np.isnan(out).sum())), flush=True)
  if stage == "fwd":
    if rel(out, ref.detach().numpy()) > 1e-3:
      r = ref.detach().numpy()
      bad = np.abs(out - r).reshape(bsz, lq // 128, 128, dv).max(axis=(2, 3))
      print("   per (image, query tile) max abs err:", np.round(bad, 3).tolist()[:2], "col err:", np.round(np.abs(out - r).max(axis=(0, 1))[:dv:8], 3).tolist())
    return
  gyd = K.from_numpy(gy)
  gyd.tf32 = True
  grads = tape.backward([(y, gyd)], [td, pd, gd], K.add)
  torch.cuda.synchronize()
  for name, a, r in zip(("dq", "dk", "dv"), grads, (tt.grad, pt.grad, gt.grad)):
    print("      %s err %.2e (nan %d)" % (name, rel(a.cpu(), r.numpy()), int(np.isnan(a.cpu()).sum())), flush=True)


def timeit(fn, iters=10):
  fn(); fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(iters):
    fn()
  e1.record()
  torch.cuda.synchronize()
  return e0.elapsed_time(e1) / iters


def timing(bsz, lq, lk, dk, dv):
  dev = K._RT["device"]
  th, ph, g, gy = [tape.DT(torch.randn(*s, device=dev) * 0.3) for s in ((bsz, lq, dk), (bsz, lk, dk), (bsz, lk, dv), (bsz, lq, dv))]
  out, lse = K.empty(bsz, lq, dv), K.empty(bsz, lq)
  dq, dkk, dvv = K.empty(bsz, lq, dk), K.empty(bsz, lk, dk), K.empty(bsz, lk, dv)
  f = lambda: K._call("attention_fwd", th.ptr, ph.ptr, g.ptr, out.ptr, lse.ptr, bsz, lq, lk, dk, dv)
  b = lambda: K._call("attention_bwd", th.ptr, ph.ptr, g.ptr, out.ptr, lse.ptr, gy.ptr, dq.ptr, dkk.ptr, dvv.ptr, bsz, lq, lk, dk, dv)
  tf, tb = timeit(f), timeit(b)
  flop_f = 2.0 * bsz * lq * lk * (2 * dk + dv)
  flop_b = 2.0 * bsz * lq * lk * (2 * (dk + dv) + 2 * dk + dv + dk)
  print("fused    B=%d lq=%d lk=%d dk=%d dv=%d: fwd %.3f ms (%.0f TF/s)  bwd %.3f ms (%.0f TF/s)" %
        (bsz, lq, lk, dk, dv, tf, flop_f / tf / 1e9, tb, flop_b / tb / 1e9), flush=True)
  with tape.no_record():
    comp = lambda: K.bmm(K.softmax(K.bmm(th, ph, False, True)), g)
    try:
      tc = timeit(comp, 3)
      print("composed forward (bmm -> softmax -> bmm, scores in HBM): %.3f ms" % tc, flush=True)
    except Exception as e:      # out of memory at the largest batch
      print("composed path failed:", str(e)[:100])


if __name__ == "__main__":
  stage = sys.argv[1] if len(sys.argv) > 1 else "fwd"
  K.init(0)
  K.set_math_mode(1)
  if stage in ("fwd", "bwd"):
    for case in [(2, 256, 128, 24, 96), (1, 128, 128, 32, 128), (2, 256, 128, 4, 16), (2, 1024, 256, 12, 48), (2, 4096, 1024, 24, 96)]:
      check(stage, *case)
  else:
    timing(64, 4096, 1024, 24, 96)
    timing(256, 4096, 1024, 24, 96)
    timing(512, 4096, 1024, 12, 48)
