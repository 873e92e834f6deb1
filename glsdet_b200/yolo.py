"""Drop-in for yolox-drone/yolo.py: the `YOLO` facade on the native path.

Same constructor keywords / `_defaults`, same methods (`generate` :99-111, `detect_image` :116-193, `get_FPS` :195-243,
`get_map_txt` :245-307) and the same results; what runs where:

  host   cvtColor, drawing (PIL), yolo_correct_boxes (numpy) and the detection-results text format - as in the reference;
  device resize_image (Pillow's BICUBIC resampler + letterbox, bit-exact, csrc/resize.cu), preprocess_input + transpose
         (fused into the Focus kernel of the backbone), backbone, neck, head, decode, score filter and NMS.

`config_path` accepts the reference's own module paths ('models/ffa/yolox_ffa.py', 'models/new/yolox10.py', ...) and maps
them to the native modules of this package, or a path / dotted name of a native module.  `state_dict=` may be given
instead of `model_path` (tests, random-init benchmarks).  With decode_mode='default' the fused detect path is used
(YoloBody.detect_uint8); the other decode modes (yolo.py:75-82) run net(images) -> decode_func -> non_max_suppression on
the device exactly like the reference's call sequence.
"""
from __future__ import annotations

import colorsys
import importlib
import os
import time
from typing import List, Optional

import numpy as np
import torch
import torch.nn as nn

from .utils import cvtColor, get_classes, preprocess_input, resize_image
from .utils_bbox import (decode_outputs, decode_outputs_cls_sigmoid, decode_outputs_no_sigmoid, decode_outputs_no_sigmoid_all,
                         detection_lines, non_max_suppression, yolo_correct_boxes)

# the reference's model files -> native modules of this package
CONFIG_MODULES = {
    "models.ffa.yolox_ffa": "glsdet_b200.yolox_ffa",
    "models.new.yolox10": "glsdet_b200.yolox10",
    "models.block.non_local.yolo_patch_nonlocal_plus": "glsdet_b200.yolo_patch_nonlocal_plus",
    "models.base.yolox": "glsdet_b200.yolox_base",
}


class YOLO(object):
    _defaults = {
        "model_path": "model_data/yolox_s.pth",
        "config_path": "models/ffa/yolox_ffa.py",
        "decode_mode": "default",
        "classes_path": "model_data/visdrone10.txt",
        "input_shape": [768, 768],        # multiples of 32 (yolo.py:34-36)
        "phi": "s",
        "confidence": 0.01,               # yolo.py:44
        "nms_iou": 0.65,                  # yolo.py:48
        "letterbox_image": False,         # yolo.py:53
        "cuda": True,
    }

    @classmethod
    def get_defaults(cls, n):
        if n in cls._defaults:
            return cls._defaults[n]
        return "Unrecognized attribute name '" + n + "'"

    def __init__(self, **kwargs):
        self.__dict__.update(self._defaults)
        self.state_dict = None
        self.class_names = None
        for name, value in kwargs.items():
            setattr(self, name, value)
        mode = self.decode_mode.strip()
        self.decode_func = {"default": decode_outputs, "obj_sigmoid": decode_outputs_no_sigmoid,
                            "no_sigmoid": decode_outputs_no_sigmoid_all, "cls_sigmoid": decode_outputs_cls_sigmoid}.get(mode)
        if self.decode_func is None:
            raise ValueError(f"unknown decode_mode {self.decode_mode!r}")
        if self.class_names is None:
            self.class_names, self.num_classes = get_classes(self.classes_path)
        else:
            self.class_names = list(self.class_names)
            self.num_classes = len(self.class_names)
        hsv_tuples = [(x / self.num_classes, 1., 1.) for x in range(self.num_classes)]
        self.colors = list(map(lambda x: colorsys.hsv_to_rgb(*x), hsv_tuples))
        self.colors = list(map(lambda x: (int(x[0] * 255), int(x[1] * 255), int(x[2] * 255)), self.colors))
        self.last_results = None
        self.generate()

    # ------------------------------------------------------------------ model
    def _module_name(self) -> str:
        path = self.config_path
        name = path[:-3].replace("/", ".") if path.endswith(".py") else path
        return CONFIG_MODULES.get(name, name)

    def generate(self):
        """yolo.py:99-111: import the model module, strict load, eval, DataParallel + cuda."""
        x = importlib.import_module(self._module_name())
        self.net = x.YoloBody(self.num_classes, self.phi)
        if not (self.cuda and torch.cuda.is_available()):
            raise RuntimeError("glsdet_b200.yolo.YOLO runs on a CUDA device: the native path has no CPU fallback")
        device = torch.device("cuda")
        sd = self.state_dict if self.state_dict is not None else torch.load(self.model_path, map_location=device)
        self.net.load_state_dict(sd)
        self.net = self.net.eval()
        self.net = nn.DataParallel(self.net)
        self.net = self.net.cuda()
        self._body = self.net.module

    # ------------------------------------------------------------------ inference
    def _input_u8(self, image) -> torch.Tensor:
        """cvtColor'ed PIL image -> uint8 [1, H, W, 3] on the device (yolo.py:130: resize_image with BICUBIC / letterbox)."""
        canvas = resize_image(image, (self.input_shape[1], self.input_shape[0]), self.letterbox_image)
        return canvas.unsqueeze(0)

    def _correct(self, rows: np.ndarray, image_shape) -> np.ndarray:
        """utils_bbox.py:478-483 on one image's NMS rows."""
        out = rows.copy()
        box_xy, box_wh = (out[:, 0:2] + out[:, 2:4]) / 2, out[:, 2:4] - out[:, 0:2]
        out[:, :4] = yolo_correct_boxes(box_xy, box_wh, self.input_shape, image_shape, self.letterbox_image)
        return out

    @torch.no_grad()
    def _infer(self, images_u8: torch.Tensor, image_shapes) -> List[Optional[np.ndarray]]:
        """uint8 [B, H, W, 3] device batch -> per image the rows non_max_suppression returns (utils_bbox.py:375-484)."""
        fused = self.decode_mode.strip() == "default" and hasattr(self._body, "detect_uint8")
        if fused:
            try:
                det, cnt = self._body.detect_uint8(images_u8, conf_thres=self.confidence, nms_thres=self.nms_iou)
            except NotImplementedError:
                fused = False
        if fused:
            counts = cnt.cpu().numpy()
            kmax = int(counts.max())
            rows = det[:, :kmax].cpu().numpy()
            return [self._correct(rows[i, :counts[i]], image_shapes[i]) for i in range(images_u8.shape[0])]
        # the reference's own call sequence (yolo.py:134-149): host preprocessing, net(images), decode_func, NMS
        results = []
        for i in range(images_u8.shape[0]):
            image_data = np.expand_dims(np.transpose(preprocess_input(images_u8[i].cpu().numpy().astype("float32")), (2, 0, 1)), 0)
            images = torch.from_numpy(image_data).cuda()
            outputs = self.net(images)
            outputs = self.decode_func(outputs, self.input_shape)
            results += non_max_suppression(outputs, self.num_classes, self.input_shape, image_shapes[i], self.letterbox_image,
                                           conf_thres=self.confidence, nms_thres=self.nms_iou)
        return results

    def detect(self, image) -> Optional[np.ndarray]:
        """The rows yolo.py works with (results[0] of :145): [K, 7] = (top, left, bottom, right, obj_conf, class_conf,
        class_pred) in pixels of `image`."""
        image_shape = np.array(np.shape(image)[0:2])
        image = cvtColor(image)
        self.last_results = self._infer(self._input_u8(image), [image_shape])
        return self.last_results[0]

    def detect_batch(self, images) -> List[Optional[np.ndarray]]:
        """Several images in one pass of the native path (not in the reference, which loops over images)."""
        shapes = [np.array(np.shape(im)[0:2]) for im in images]
        batch = torch.cat([self._input_u8(cvtColor(im)) for im in images])
        self.last_results = self._infer(batch, shapes)
        return self.last_results

    def detect_image(self, image):
        """yolo.py:116-193: returns the PIL image with the detections drawn."""
        from PIL import ImageDraw, ImageFont

        rows = self.detect(image)
        image = cvtColor(image)
        if rows is None:
            return image
        top_label = np.array(rows[:, 6], dtype="int32")
        top_conf = rows[:, 4] * rows[:, 5]
        top_boxes = rows[:, :4]
        size = int(np.floor(3e-2 * image.size[1] + 0.5).astype("int32"))
        try:
            font = ImageFont.truetype(font="model_data/simhei.ttf", size=size)
        except OSError:
            font = ImageFont.load_default()
        thickness = int(max((image.size[0] + image.size[1]) // np.mean(self.input_shape), 1))
        for i, c in list(enumerate(top_label)):
            predicted_class = self.class_names[int(c)]
            top, left, bottom, right = top_boxes[i]
            top = max(0, np.floor(top).astype("int32"))
            left = max(0, np.floor(left).astype("int32"))
            bottom = min(image.size[1], np.floor(bottom).astype("int32"))
            right = min(image.size[0], np.floor(right).astype("int32"))
            label = "{} {:.2f}".format(predicted_class, top_conf[i])
            draw = ImageDraw.Draw(image)
            if hasattr(draw, "textsize"):
                label_size = draw.textsize(label, font)
            else:   # Pillow >= 10
                l, t, r, b = draw.textbbox((0, 0), label, font=font)
                label_size = (r - l, b - t)
            if top - label_size[1] >= 0:
                text_origin = np.array([left, top - label_size[1]])
            else:
                text_origin = np.array([left, top + 1])
            for k in range(thickness):
                if right - k >= left + k and bottom - k >= top + k:
                    draw.rectangle([left + k, top + k, right - k, bottom - k], outline=self.colors[c])
            draw.rectangle([tuple(text_origin), tuple(text_origin + label_size)], fill=self.colors[c])
            draw.text(tuple(text_origin), label, fill=(0, 0, 0), font=font)
            del draw
        return image

    def get_FPS(self, image, test_interval):
        """yolo.py:195-243: seconds per image of the device part (upload excluded, like the reference which times
        net + decode + NMS on an already uploaded tensor)."""
        image_shape = np.array(np.shape(image)[0:2])
        image = cvtColor(image)
        batch = self._input_u8(image)
        self._infer(batch, [image_shape])
        torch.cuda.synchronize()
        t1 = time.time()
        for _ in range(test_interval):
            self._infer(batch, [image_shape])
        torch.cuda.synchronize()
        t2 = time.time()
        return (t2 - t1) / test_interval

    def get_map_txt(self, image_id, image, class_names, map_out_path):
        """yolo.py:245-307: detection-results/<image_id>.txt, one "<class> <score[:6]> <left> <top> <right> <bottom>" line
        per detection of a class in `class_names`."""
        os.makedirs(os.path.join(map_out_path, "detection-results"), exist_ok=True)
        with open(os.path.join(map_out_path, "detection-results/" + image_id + ".txt"), "w") as f:
            rows = self.detect(image)
            if rows is None:
                return
            for line in detection_lines(rows, self.class_names, class_names):
                f.write(line + "\n")
        return
