from video_diffusion import _extend_with_reference  # noqa: E402

_extend_with_reference(__path__, __file__, "models")
