class AutoencoderKL:  # placeholder
    pass
