// rsv_lines_abi.inl — C ABI of the batched LineTest (added with the DDA kernel).
