"""pypanadapter_b200 -- B200-native zoom-FFT PSD path behind pypanadapter's call surface."""
