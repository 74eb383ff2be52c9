"""EuRoC MAV on-disk format: reader (drop-in for the reference's src/streaming/dataset.py:189-220) and writer.

Layout (dataset.py:193-220):
    <seq>/mav0/cam{0,1}/data/<t_ns>.png                     8-bit grayscale, file name = timestamp in ns
    <seq>/mav0/imu0/data.csv                                header line, then t_ns,wx,wy,wz,ax,ay,az
    <seq>/mav0/state_groundtruth_estimate0/data.csv         header line, then t_ns,p(3),q_wxyz(4),v(3),bw(3),ba(3)
Start time = max(first IMU stamp, first image stamp) + offset (dataset.py:206-214).

The reference decodes with cv2.imread; here the PNG container is parsed in Python (zlib inflate) and the scanline
filters are undone in C (_avbhost.png_unfilter), optionally straight into a caller-provided buffer such as libavb's
pinned staging block, so a frame goes disk -> pinned -> HBM with no intermediate copy.  `events()` adds the
deterministic IMU/stereo interleave the test and bench drivers use."""
from __future__ import annotations

import os
import struct
import zlib
from collections import namedtuple

import numpy as np

from image_processing import _avbhost

imu_msg = namedtuple('imu_msg', ['timestamp', 'angular_velocity', 'linear_acceleration'])
img_msg = namedtuple('img_msg', ['timestamp', 'image'])
stereo_msg = namedtuple('stereo_msg', ['timestamp', 'cam0_image', 'cam1_image', 'cam0_msg', 'cam1_msg'])
gt_msg = namedtuple('gt_msg', ['timestamp', 'p', 'q', 'v', 'bw', 'ba'])

_PNG_SIG = b'\x89PNG\r\n\x1a\n'


# ---- PNG codec (8/16-bit grayscale, non-interlaced; what EuRoC uses) ----------------------------------------
def png_decode(data: bytes, out=None) -> np.ndarray:
    if data[:8] != _PNG_SIG:
        raise ValueError('not a PNG file')
    pos, idat, hdr = 8, [], None
    while pos < len(data):
        n, kind = struct.unpack('>I4s', data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        pos += 12 + n
        if kind == b'IHDR':
            hdr = struct.unpack('>IIBBBBB', body)
        elif kind == b'IDAT':
            idat.append(body)
        elif kind == b'IEND':
            break
    if hdr is None:
        raise ValueError('PNG without IHDR')
    w, h, depth, ctype, _, _, interlace = hdr
    if ctype != 0 or depth not in (8, 16) or interlace != 0:
        raise ValueError(f'unsupported PNG (colour type {ctype}, depth {depth}, interlace {interlace}): EuRoC frames are '
                         '8-bit grayscale')
    bpp = depth // 8
    raw = zlib.decompress(b''.join(idat))
    if out is None:
        out = np.empty((h, w * bpp), np.uint8)
    elif out.dtype != np.uint8 or out.size != h * w * bpp or not out.flags.c_contiguous:
        raise ValueError('output buffer must be a C-contiguous uint8 array of the image size')
    _avbhost.png_unfilter(raw, w, h, bpp, out)
    if bpp == 2:
        return out.reshape(h, w, 2).view('>u2').reshape(h, w)
    return out.reshape(h, w)


def png_encode(img: np.ndarray, level: int = 1) -> bytes:
    img = np.ascontiguousarray(img, dtype=np.uint8)
    if img.ndim != 2:
        raise ValueError('8-bit grayscale images only')
    h, w = img.shape
    rows = np.empty((h, w + 1), np.uint8)
    rows[:, 0] = 0                                   # filter type None on every scanline
    rows[:, 1:] = img

    def chunk(kind, body):
        return struct.pack('>I', len(body)) + kind + body + struct.pack('>I', zlib.crc32(kind + body) & 0xffffffff)

    return (_PNG_SIG + chunk(b'IHDR', struct.pack('>IIBBBBB', w, h, 8, 0, 0, 0, 0))
            + chunk(b'IDAT', zlib.compress(rows.tobytes(), level)) + chunk(b'IEND', b''))


# ---- readers with the reference's surface ----------------------------------------------------------------------
class _CsvReader:
    def __init__(self, path, scaler, starttime=-float('inf')):
        self.path, self.scaler, self.starttime = path, scaler, starttime

    def set_starttime(self, starttime):
        self.starttime = starttime

    def _rows(self):
        with open(self.path, 'r') as f:
            next(f)
            for line in f:
                if line.strip():
                    yield [float(v) for v in line.strip().split(',')]

    def start_time(self):
        for r in self._rows():
            return r[0] * self.scaler


class IMUDataReader(_CsvReader):
    def __iter__(self):
        for r in self._rows():
            t = r[0] * self.scaler
            if t >= self.starttime:
                yield imu_msg(t, np.array(r[1:4]), np.array(r[4:7]))


class GroundTruthReader(_CsvReader):
    def __iter__(self):
        for r in self._rows():
            t = r[0] * self.scaler
            if t >= self.starttime:
                yield gt_msg(t, np.array(r[1:4]), np.array(r[4:8]), np.array(r[8:11]), np.array(r[11:14]), np.array(r[14:17]))


class ImageReader:
    def __init__(self, ids, timestamps, starttime=-float('inf')):
        self.ids, self.timestamps, self.starttime = ids, timestamps, starttime

    def read(self, path, out=None):
        with open(path, 'rb') as f:
            return png_decode(f.read(), out)

    def __len__(self):
        return len(self.ids)

    def __getitem__(self, idx):
        return self.read(self.ids[idx])

    def __iter__(self):
        for i, t in enumerate(self.timestamps):
            if t >= self.starttime:
                yield img_msg(t, self[i])

    def start_time(self):
        return self.timestamps[0]

    def set_starttime(self, starttime):
        self.starttime = starttime


class Stereo:
    def __init__(self, cam0, cam1):
        self.cam0, self.cam1, self.timestamps = cam0, cam1, cam0.timestamps

    def __iter__(self):
        for l, r in zip(self.cam0, self.cam1):
            yield stereo_msg(l.timestamp, l.image, r.image, l, r)

    def __len__(self):
        return len(self.cam0)

    def prefetch(self, threads=8, depth=None):
        """Same messages as iteration, decoded ahead by `threads` worker threads (file read, inflate and the C unfilter
        all release the GIL): what fills a sweep's HBM frame store, where every frame of a sequence is decoded exactly
        once and decode speed is the set-up time."""
        from concurrent.futures import ThreadPoolExecutor
        idx = [i for i, t in enumerate(self.cam0.timestamps) if t >= self.cam0.starttime]
        depth = depth or 4 * threads

        def load(i):
            t = self.cam0.timestamps[i]
            l, r = img_msg(t, self.cam0[i]), img_msg(self.cam1.timestamps[i], self.cam1[i])
            return stereo_msg(t, l.image, r.image, l, r)

        with ThreadPoolExecutor(max_workers=threads) as pool:
            pending = []
            for i in idx:
                pending.append(pool.submit(load, i))
                if len(pending) >= depth:
                    yield pending.pop(0).result()
            for f in pending:
                yield f.result()

    def start_time(self):
        return self.cam0.starttime

    def set_starttime(self, starttime):
        self.starttime = starttime
        self.cam0.set_starttime(starttime)
        self.cam1.set_starttime(starttime)


class EuRoCDataset:
    """path example: '.../MH_01_easy'.  Same attributes as the reference's EuRoCDataset: groundtruth, imu, cam0, cam1,
    stereo, timestamps, starttime, set_starttime(offset)."""

    def __init__(self, path):
        self.path = path
        mav = os.path.join(path, 'mav0')
        self.groundtruth = GroundTruthReader(os.path.join(mav, 'state_groundtruth_estimate0', 'data.csv'), 1e-9)
        self.imu = IMUDataReader(os.path.join(mav, 'imu0', 'data.csv'), 1e-9)
        self.cam0 = ImageReader(*self.list_imgs(os.path.join(mav, 'cam0', 'data')))
        self.cam1 = ImageReader(*self.list_imgs(os.path.join(mav, 'cam1', 'data')))
        self.stereo = Stereo(self.cam0, self.cam1)
        self.timestamps = self.cam0.timestamps
        # The reference takes max(imu.start_time(), stereo.start_time()), and Stereo.start_time() returns
        # cam0.starttime, which is still -inf here (dataset.py:184-185, 203): the start time IS the first IMU stamp,
        # also when the cameras start later.  Reproduced, not repaired: every --offset run counts from there.
        self.starttime = max(self.imu.start_time(), self.stereo.start_time())
        self.set_starttime(0)

    def set_starttime(self, offset):
        t = self.starttime + offset
        for r in (self.groundtruth, self.imu, self.cam0, self.cam1, self.stereo):
            r.set_starttime(t)

    @staticmethod
    def list_imgs(d):
        xs = sorted((x for x in os.listdir(d) if x.endswith('.png')), key=lambda x: float(x[:-4]))
        return [os.path.join(d, x) for x in xs], [float(x[:-4]) * 1e-9 for x in xs]

    def events(self):
        """('imu', msg) / ('stereo', msg): before each stereo frame every IMU message with timestamp <= the frame's."""
        imu_it = iter(self.imu)
        pending = next(imu_it, None)
        for f in self.stereo:
            while pending is not None and pending.timestamp <= f.timestamp:
                yield 'imu', pending
                pending = next(imu_it, None)
            yield 'stereo', f


# ---- writer ----------------------------------------------------------------------------------------------------
def write_euroc(path, stream, groundtruth=None):
    """Writes a stream (objects with .frames(), .imu() and optionally .groundtruth()) in the EuRoC layout."""
    mav = os.path.join(path, 'mav0')
    for sub in ('cam0/data', 'cam1/data', 'imu0', 'state_groundtruth_estimate0'):
        os.makedirs(os.path.join(mav, sub), exist_ok=True)
    for f in stream.frames():
        name = '%d.png' % int(round(f.timestamp * 1e9))
        for cam, img in (('cam0', f.cam0_image), ('cam1', f.cam1_image)):
            with open(os.path.join(mav, cam, 'data', name), 'wb') as fh:
                fh.write(png_encode(img))
    with open(os.path.join(mav, 'imu0', 'data.csv'), 'w') as fh:
        fh.write('#timestamp [ns],w_RS_S_x [rad s^-1],w_RS_S_y [rad s^-1],w_RS_S_z [rad s^-1],'
                 'a_RS_S_x [m s^-2],a_RS_S_y [m s^-2],a_RS_S_z [m s^-2]\n')
        for m in stream.imu():
            fh.write('%d,' % int(round(m.timestamp * 1e9)) + ','.join(repr(float(v)) for v in
                                                                      list(m.angular_velocity) + list(m.linear_acceleration)) + '\n')
    gt = groundtruth if groundtruth is not None else (stream.groundtruth() if hasattr(stream, 'groundtruth') else [])
    with open(os.path.join(mav, 'state_groundtruth_estimate0', 'data.csv'), 'w') as fh:
        fh.write('#timestamp,p_RS_R_x [m],p_RS_R_y [m],p_RS_R_z [m],q_RS_w [],q_RS_x [],q_RS_y [],q_RS_z [],'
                 'v_RS_R_x [m s^-1],v_RS_R_y [m s^-1],v_RS_R_z [m s^-1],b_w_RS_S_x [rad s^-1],b_w_RS_S_y [rad s^-1],'
                 'b_w_RS_S_z [rad s^-1],b_a_RS_S_x [m s^-2],b_a_RS_S_y [m s^-2],b_a_RS_S_z [m s^-2]\n')
        for g in gt:
            vals = list(g.p) + list(g.q) + list(g.v) + list(g.bw) + list(g.ba)
            fh.write('%d,' % int(round(g.timestamp * 1e9)) + ','.join(repr(float(v)) for v in vals) + '\n')
