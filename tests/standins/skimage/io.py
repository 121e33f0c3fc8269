import cv2
import numpy as np


def imread(path, *a, **k):
    img = cv2.imread(str(path), cv2.IMREAD_UNCHANGED)
    if img is None:
        raise FileNotFoundError(path)
    if img.ndim == 3 and img.shape[2] >= 3:
        img = img[:, :, [2, 1, 0] + list(range(3, img.shape[2]))]        # BGR(A) -> RGB(A)
    return np.ascontiguousarray(img)


def imsave(path, arr, *a, **k):
    arr = np.asarray(arr)
    if arr.ndim == 3 and arr.shape[2] >= 3:
        arr = arr[:, :, [2, 1, 0] + list(range(3, arr.shape[2]))]
    cv2.imwrite(str(path), arr)
