"""Measurement tools of the reference that sit on top of the inference path (src/neural_audio_spring_reverb/tools/)."""
