{
  "targets": [{
    "target_name": "nzcb_napi",
    "sources": ["nzcb_napi.c"],
    "include_dirs": ["../include"],
    "libraries": ["-L<(module_root_dir)/../nzcb_circom_b200", "-lnzcb", "-Wl,-rpath,<(module_root_dir)/../nzcb_circom_b200"]
  }]
}
