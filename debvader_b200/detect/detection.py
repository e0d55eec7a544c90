"""reference detect/detection.py:5-56 — SExtractor detection through the third-party `sep`
C library.  Detection is OUT OF SCOPE of the B200 hot path (centres are an input to it); this is
the same call sequence with a lazy import so the package works without `sep` installed."""
import numpy as np

# 7x7 convolution mask of a gaussian PSF with FWHM = 3.0 pixels (detection.py:25-35)
FILTER_KERNEL = np.array(
    [
        [0.004963, 0.021388, 0.051328, 0.068707, 0.051328, 0.021388, 0.004963],
        [0.021388, 0.092163, 0.221178, 0.296069, 0.221178, 0.092163, 0.021388],
        [0.051328, 0.221178, 0.530797, 0.710525, 0.530797, 0.221178, 0.051328],
        [0.068707, 0.296069, 0.710525, 0.951108, 0.710525, 0.296069, 0.068707],
        [0.051328, 0.221178, 0.530797, 0.710525, 0.530797, 0.221178, 0.051328],
        [0.021388, 0.092163, 0.221178, 0.296069, 0.221178, 0.092163, 0.021388],
        [0.004963, 0.021388, 0.051328, 0.068707, 0.051328, 0.021388, 0.004963],
    ]
)


def detect_objects(field_image):
    """Detect objects on the r band (index 2) with sep; returns (row, col) offsets from the centre."""
    try:
        import sep
    except ImportError as e:  # pragma: no cover
        raise ImportError("detect_objects needs the third-party `sep` package (SExtractor); "
                          "pass galaxy_distances_to_center / a detector= callable instead") from e
    if hasattr(field_image, "detach"):
        field_image = field_image.detach().cpu().numpy()
    field_image = np.asarray(field_image).copy()
    field_size = field_image.shape[1]
    r_band = field_image[0, :, :, 2].copy()
    bkg = sep.Background(r_band)
    objects = sep.extract(data=r_band - bkg, thresh=1.5, err=bkg.globalrms, deblend_cont=0.00001, deblend_nthresh=64,
                          minarea=4, filter_kernel=FILTER_KERNEL, filter_type="conv")
    out = [(np.round(-int(field_size / 2) + objects["y"][i]), np.round(-int(field_size / 2) + objects["x"][i]))
           for i in range(len(objects["y"]))]
    return np.array(out)
