"""Import stub (skimage is not installed in this image); see pytorch_lightning stub."""
