"""Default bounding boxes (anchors) - host-side mirror of the reference's `ssdseglib/boxes.py`.

Same public surface as the reference (`DefaultBoundingBoxes`, its ten getters, the two free
corner<->centroid converters; reference boxes.py:5-445).  Anchor construction is a one-time host step in
the reference (1.7 ms for 9600 boxes) and stays on the host here: the device path starts where the
anchors are uploaded (`datacoder.DataEncoderDecoder`, `layers.DecodeBoxesCentroidsOffsets`).

Numerics follow the reference so the anchor table is bit-identical: float64 arithmetic, one cast to
float32 when the corners are stored (boxes.py:139-146), an fp32 multiply for the rescale (boxes.py:167-174).
"""
from __future__ import annotations

import math
from typing import List, Literal, Sequence, Tuple, Union

import numpy as np

Number = Union[int, float]
_CORNER_SLOT = {"xmin": 0, "ymin": 1, "xmax": 2, "ymax": 3}
_CENTROID_SLOT = {"center-x": 0, "center-y": 1, "width": 2, "height": 3}


def _is_number(v) -> bool:
    return isinstance(v, (int, float))


class DefaultBoundingBoxes:
    """SSD default boxes for a set of feature maps (reference boxes.py:5-388).

    Attributes kept from the reference: `feature_maps_shapes`, `feature_maps_aspect_ratios`,
    `additional_square_box`, `boxes_scales`, `centers_padding_from_borders_percentage`,
    `feature_maps_boxes` (None until `rescale_boxes_coordinates` is called).
    """

    def __init__(
        self,
        feature_maps_shapes: Tuple[Tuple[int, int], ...],
        feature_maps_aspect_ratios: Union[Tuple[Number, ...], Tuple[Tuple[Number, ...], ...]] = (1, 2, 3, 1 / 2, 1 / 3),
        boxes_scales: Tuple[float, float] = (0.2, 0.9),
        centers_padding_from_borders_percentage: Union[float, Tuple[float, ...]] = 0.05,
        additional_square_box: bool = True,
    ) -> None:
        self.feature_maps_shapes = feature_maps_shapes
        self.additional_square_box = additional_square_box
        n_maps = len(feature_maps_shapes)

        # padding: one float for every map, or one float per map, each in [0, 0.5)  (boxes.py:39-44)
        pad = centers_padding_from_borders_percentage
        if isinstance(pad, float) and 0 <= pad < 0.5:
            self.centers_padding_from_borders_percentage = (pad,) * n_maps
        elif (not isinstance(pad, (int, float)) and len(pad) == n_maps
              and all(isinstance(p, float) and 0 <= p < 0.5 for p in pad)):
            self.centers_padding_from_borders_percentage = pad
        else:
            raise ValueError('the percentage padding from borders should be a float or a tuple of floats, '
                             'with values in the range [0, 0.5)')

        # one extra scale feeds the additional square box of the last map (boxes.py:47)
        self.boxes_scales = np.linspace(boxes_scales[0], boxes_scales[1], n_maps + 1)

        # aspect ratios: shared tuple of numbers, or one tuple per map (boxes.py:51-58)
        if all(_is_number(r) for r in feature_maps_aspect_ratios):
            self.feature_maps_aspect_ratios = tuple(feature_maps_aspect_ratios for _ in range(n_maps))
        elif len(feature_maps_aspect_ratios) < n_maps:
            raise ValueError('if you are passing a tuple of tuples of aspect ratios, then it should have same '
                             'length as the tuple of feature maps shapes')
        else:
            self.feature_maps_aspect_ratios = feature_maps_aspect_ratios

        self._feature_maps_boxes = self._generate_feature_maps_boxes()
        self.feature_maps_boxes = None

    # ------------------------------------------------------------------ construction
    @staticmethod
    def _axis_centers(n: int, padding: float) -> np.ndarray:
        """Box centres along one axis of an n-pixel map, in pixel-index units (boxes.py:112-130)."""
        if n == 1:
            return np.array([0.5])
        margin = padding * (n - 1.0)
        return np.linspace(start=margin, stop=n - 1.0 - margin, num=n)

    def _generate_feature_maps_boxes(self) -> List[np.ndarray]:
        """One (H, W, boxes, 4) float32 array per feature map, corners (xmin, ymin, xmax, ymax) in
        feature-map pixel units (boxes.py:74-151)."""
        per_map = []
        for m, (shape, ratios, padding) in enumerate(
                zip(self.feature_maps_shapes, self.feature_maps_aspect_ratios,
                    self.centers_padding_from_borders_percentage)):
            rows, cols = shape
            side = min(shape)
            scale, scale_next = self.boxes_scales[m], self.boxes_scales[m + 1]
            heights = [side * scale / math.sqrt(r) for r in ratios]
            widths = [side * scale * math.sqrt(r) for r in ratios]
            if self.additional_square_box:
                extra = side * math.sqrt(scale * scale_next)
                heights.append(extra)
                widths.append(extra)
            half_h = (np.array(heights) - 1.0) / 2.0
            half_w = (np.array(widths) - 1.0) / 2.0
            cy = self._axis_centers(rows, padding)[:, None, None]
            cx = self._axis_centers(cols, padding)[None, :, None]
            boxes = np.zeros((rows, cols, len(heights), 4), dtype=np.float32)
            boxes[..., 0] = cx - half_w
            boxes[..., 1] = cy - half_h
            boxes[..., 2] = cx + half_w
            boxes[..., 3] = cy + half_h
            per_map.append(boxes)
        return per_map

    def rescale_boxes_coordinates(self, image_shape: Tuple[int, int]) -> None:
        """Scale the boxes to an image of `image_shape` = (height, width) (boxes.py:153-177).

        Reference quirk kept on purpose: the internal arrays are scaled IN PLACE and then aliased as
        `feature_maps_boxes`, so a second call scales the already scaled boxes again."""
        self.feature_maps_boxes = []
        for boxes, (rows, cols) in zip(self._feature_maps_boxes, self.feature_maps_shapes):
            fx = (image_shape[1] - 1) / (cols - 1 if cols > 1 else 1)
            fy = (image_shape[0] - 1) / (rows - 1 if rows > 1 else 1)
            boxes[..., [0, 2]] = boxes[..., [0, 2]] * fx
            boxes[..., [1, 3]] = boxes[..., [1, 3]] * fy
            self.feature_maps_boxes.append(boxes)

    # ------------------------------------------------------------------ getters
    @staticmethod
    def _select(per_map: Sequence[np.ndarray], slot, style: str):
        """Common tail of every getter: pick coordinate column(s) and apply the output style."""
        cols = [0, 1, 2, 3] if slot is None else (slot,)
        picked = tuple(b[:, :, :, cols] for b in per_map)
        if style == 'ssd':
            flat_shape = (-1, 4) if slot is None else (-1,)
            return np.concatenate([np.reshape(p, flat_shape) for p in picked], axis=0)
        return picked

    def _get_boxes_coordinates_corners_all_or_single(
            self, coordinates_type: Literal['xmin', 'ymin', 'xmax', 'ymax', 'corners'],
            coordinates_style: Literal['ssd', 'feature-maps']):
        slot = None if coordinates_type == 'corners' else _CORNER_SLOT[coordinates_type]
        return self._select(self.feature_maps_boxes, slot, coordinates_style)

    def _get_boxes_coordinates_centroids_all_or_single(
            self, coordinates_type: Literal['center-x', 'center-y', 'width', 'height', 'centroids'],
            coordinates_style: Literal['ssd', 'feature-maps']):
        per_map = []
        for corners in self.feature_maps_boxes:  # fp32 arithmetic on the stored corners (boxes.py:301-310)
            x0, y0, x1, y1 = (corners[..., i:i + 1] for i in range(4))
            per_map.append(np.concatenate(coordinates_corners_to_centroids(x0, y0, x1, y1), axis=3))
        slot = None if coordinates_type == 'centroids' else _CENTROID_SLOT[coordinates_type]
        return self._select(per_map, slot, coordinates_style)

    def get_boxes_coordinates_corners(self, coordinates_style: Literal['ssd', 'feature-maps']):
        """All boxes as corners: (A, 4) for 'ssd', a tuple of (H, W, boxes, 4) for 'feature-maps'."""
        return self._get_boxes_coordinates_corners_all_or_single('corners', coordinates_style)

    def get_boxes_coordinates_xmin(self, coordinates_style: Literal['ssd', 'feature-maps']):
        return self._get_boxes_coordinates_corners_all_or_single('xmin', coordinates_style)

    def get_boxes_coordinates_ymin(self, coordinates_style: Literal['ssd', 'feature-maps']):
        return self._get_boxes_coordinates_corners_all_or_single('ymin', coordinates_style)

    def get_boxes_coordinates_xmax(self, coordinates_style: Literal['ssd', 'feature-maps']):
        return self._get_boxes_coordinates_corners_all_or_single('xmax', coordinates_style)

    def get_boxes_coordinates_ymax(self, coordinates_style: Literal['ssd', 'feature-maps']):
        return self._get_boxes_coordinates_corners_all_or_single('ymax', coordinates_style)

    def get_boxes_coordinates_centroids(self, coordinates_style: Literal['ssd', 'feature-maps']):
        """All boxes as centroids (center_x, center_y, width, height)."""
        return self._get_boxes_coordinates_centroids_all_or_single('centroids', coordinates_style)

    def get_boxes_coordinates_center_x(self, coordinates_style: Literal['ssd', 'feature-maps']):
        return self._get_boxes_coordinates_centroids_all_or_single('center-x', coordinates_style)

    def get_boxes_coordinates_center_y(self, coordinates_style: Literal['ssd', 'feature-maps']):
        return self._get_boxes_coordinates_centroids_all_or_single('center-y', coordinates_style)

    def get_boxes_coordinates_width(self, coordinates_style: Literal['ssd', 'feature-maps']):
        return self._get_boxes_coordinates_centroids_all_or_single('width', coordinates_style)

    def get_boxes_coordinates_height(self, coordinates_style: Literal['ssd', 'feature-maps']):
        return self._get_boxes_coordinates_centroids_all_or_single('height', coordinates_style)


def coordinates_corners_to_centroids(xmin, ymin, xmax, ymax):
    """(xmin, ymin, xmax, ymax) -> (center_x, center_y, width, height); pixel coordinates are inclusive
    indices, hence the +1 (reference boxes.py:391-417)."""
    return (xmax + xmin) / 2.0, (ymax + ymin) / 2.0, xmax - xmin + 1.0, ymax - ymin + 1.0


def coordinates_centroids_to_corners(center_x, center_y, width, height):
    """(center_x, center_y, width, height) -> (xmin, ymin, xmax, ymax) (reference boxes.py:419-445)."""
    half_w = (width - 1.0) / 2.0
    half_h = (height - 1.0) / 2.0
    return center_x - half_w, center_y - half_h, center_x + half_w, center_y + half_h
