// placeholder (filled by the batched position fit)
