// not used by the encoder sources
