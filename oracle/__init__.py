"""ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.

CPU (fp32 torch / numpy) restatement of the reference's per-frame LCM img2img path, used as the checker for the
CUDA implementation in videosd_b200/. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package; nothing under videosd_b200/ does.

Why a restatement: the arithmetic of the reference path lives in the third-party, un-vendored, UNPINNED package
`diffusers` (diffusert/requirements.txt:1, effective version ~0.23, Nov 2023) plus PyAV/libswscale for the colour
conversion; neither is installed here or on the GPU box and there is no network. The reference itself has no tests,
golden vectors or fixtures, so this oracle cannot be pinned against reference outputs:
"parity unpinned". What IS checked (tests/test_oracle.py):
  * parameter-count identities of the restated modules (UNet 859 602 884; TAESD 1 222 532 / 1 222 531),
  * diffusers state-dict key names (so real checkpoints would load),
  * scheduler constants / timestep tables computed from the reference's own formulas
    (diffusert/lcm/lcm_controlnet.py:793-815, :905-946),
  * the reference's live scheduler code itself: tests import `LCMScheduler_X` from /root/reference (in this
    container only, against a stub `diffusers` namespace) and compare step/add_noise/set_timesteps outputs with
    oracle.scheduler; the resulting vectors are committed under tests/golden/ with the generating script.

Modules:
  scheduler.py  LCMScheduler_X restatement            (lcm_controlnet.py:713-1071)
  unet.py       UNet2DConditionModel (SD1.5 + LCM)    (diffusers; call site lcm_controlnet.py:568-577)
  taesd.py      AutoencoderTiny                        (diffusers; call sites :298-300, :594-596)
  imageproc.py  VaeImageProcessor pre/post + the YUV420<->RGB integer spec (server.py:108,117)
  pipeline.py   frame sequencing and RNG order         (lcm_controlnet.py:380-618, videopipeline.py:75-128)
  weights.py    deterministic random-init recipe
"""
