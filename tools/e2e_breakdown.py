"""Where the end-to-end frame time goes on the host side (C2, one stream): device time of the frame (CUDA events),
the C driver call, the whole stereo_callback."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'uav-airvision_b200')):
    sys.path.insert(0, p)
import torch
from bench import workload
from synth_euroc import SlidingTextureStream
from image_processing import ImageProcessor, _avbhost, FeatureMeasurement
from synth_euroc import img_msg, stereo_msg

cfg, skw, _ = workload('c2')
n = 260
stream = SlidingTextureStream(n_frames=n, **skw)
pin = torch.empty((n, 2, stream.h, stream.w), dtype=torch.uint8).pin_memory().numpy()
evs = []
k = 0
for kind, msg in stream.events():
    if kind == 'stereo':
        pin[k, 0], pin[k, 1] = msg.cam0_image, msg.cam1_image
        msg = stereo_msg(msg.timestamp, pin[k, 0], pin[k, 1], img_msg(msg.timestamp, pin[k, 0]), img_msg(msg.timestamp, pin[k, 1]))
        k += 1
    evs.append((kind, msg))
ip = ImageProcessor(cfg)
t_cb, t_imu, dev = [], [], []
orig = ip.imu_processor.integrate_imu_data
def timed_imu():
    t0 = time.perf_counter(); r = orig(); t_imu.append(time.perf_counter() - t0); return r
ip.imu_processor.integrate_imu_data = timed_imu
t_pf, t_sub, t_fin = [], [], []
import image_processing.pipeline as pl


def _timed(fn, acc):
    def f(*a):
        t0 = time.perf_counter(); r = fn(*a); acc.append(time.perf_counter() - t0); return r
    return f


class _H:                               # the C driver calls of stereo_callback, timed
    process_frame = staticmethod(_timed(_avbhost.process_frame, t_pf))
    submit_images = staticmethod(_timed(_avbhost.submit_images, t_sub))
    finish_frame = staticmethod(_timed(_avbhost.finish_frame, t_fin))


pl._avbhost = _H
for kind, msg in evs:
    if kind == 'imu':
        ip.imu_callback(msg)
    else:
        t0 = time.perf_counter(); fm = ip.stereo_callback(msg); t_cb.append(time.perf_counter() - t0)
        dev.append(ip.context.last_frame_ms() if hasattr(ip.context, 'last_frame_ms') else float('nan'))
med = lambda a: 1e6 * float(np.median(a[20:]))
print(f'stereo_callback {med(t_cb):.1f} us | _avbhost.submit_images {med(t_sub):.1f} us + finish_frame {med(t_fin):.1f} us | integrate_imu {med(t_imu):.1f} us (between the two) | '
      f'device (events, H2D..results) {1e3 * float(np.median(dev[20:])):.1f} us | features {len(fm.features)}')
