"""PyramidBuilder: same constructor / attributes / return value as the reference's placeholder
(image_processing/pyramid_builder.py:5-48, which hands the raw images to cv2 and lets every calcOpticalFlowPyrLK call
rebuild the pyramids).  Here `create_image_pyramids` makes the frame current in the libavb context (the previous
frame's device pyramids are kept), uploads both images and builds levels 1..L once with k_pyr_down / k_pyr_pair."""
from __future__ import annotations


class PyramidBuilder:
    def __init__(self, win_size, pyramid_levels, cam0_curr_img_msg, cam1_curr_img_msg, context=None):
        self.win_size = win_size
        self.pyramid_levels = pyramid_levels
        self.cam0_curr_img_msg = cam0_curr_img_msg
        self.cam1_curr_img_msg = cam1_curr_img_msg
        self.curr_cam0_pyramid = None
        self.curr_cam1_pyramid = None
        self._ctx = context

    def _context(self):
        if self._ctx is None:
            from .pipeline import current_context
            self._ctx = current_context()
        return self._ctx

    def create_image_pyramids(self):
        img0, img1 = self.cam0_curr_img_msg.image, self.cam1_curr_img_msg.image
        ctx = self._context()
        if ctx.max_level != int(self.pyramid_levels):
            raise RuntimeError(f'context was built for {ctx.max_level + 1} pyramid levels, asked for {self.pyramid_levels + 1}')
        ctx.begin_frame(img0, img1)                 # roll current -> previous, upload, build the device pyramids
        # like the reference, the returned "pyramids" are the images themselves (callers only read .shape);
        # the device pyramids are addressed through the context's image slots
        self.curr_cam0_pyramid, self.curr_cam1_pyramid = img0, img1
        return img0, img1
