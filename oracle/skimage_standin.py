"""TEST INFRASTRUCTURE ONLY -- stand-in for the two skimage.measure functions the reference's get_roi calls
(reference proc/roi.py:48-49, :59-63).  scikit-image is not installed in this image, so this restates its documented
behaviour on top of scipy.ndimage: PARITY UNPINNED against skimage itself (unpinned in the reference's setup.py).

  label(bin_im)          default connectivity = input.ndim (8-connected in 2-D), background 0, regions numbered from 1 in
                         raster order of their first pixel -- scipy.ndimage.label with a full 3x3 structure numbers the
                         same way;
  regionprops(label_im)  one entry per label in increasing label order with `.area` (pixel count), `.extent`
                         (area / bounding-box area) and `.coords` ((n,2) row, col in raster order).
"""
import numpy as np
import scipy.ndimage as ndi


def label(image):
    out, _ = ndi.label(np.asarray(image) != 0, structure=np.ones((3, 3), dtype=int))
    return out.astype(np.int64)


class _Region:
    def __init__(self, lab, sl, label_im):
        self.label = lab
        rows, cols = np.nonzero(label_im[sl] == lab)
        self.coords = np.stack([rows + sl[0].start, cols + sl[1].start], axis=1)
        self.area = float(len(rows))
        self.bbox = (sl[0].start, sl[1].start, sl[0].stop, sl[1].stop)
        self.area_bbox = float((sl[0].stop - sl[0].start) * (sl[1].stop - sl[1].start))
        self.extent = self.area / self.area_bbox


def regionprops(label_im):
    label_im = np.asarray(label_im)
    return [_Region(i + 1, sl, label_im) for i, sl in enumerate(ndi.find_objects(label_im)) if sl is not None]
