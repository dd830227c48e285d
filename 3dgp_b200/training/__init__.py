"""Module surface of the hot path: tri-plane generator, fused renderer front-end, discriminator, loss and the data-parallel step."""
