"""QuantPipe on the device - same function names as the reference `pipeedge.quantization`."""
