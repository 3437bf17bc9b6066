{
  "targets": [{
    "target_name": "planet_b200_addon",
    "sources": ["planet_b200_addon.cc"],
    "include_dirs": ["../../include"],
    "libraries": ["-L<(module_root_dir)/../../planet_heightmap_generation_b200", "-lplanet_b200",
                  "-Wl,-rpath,<(module_root_dir)/../../planet_heightmap_generation_b200"],
    "cflags_cc": ["-std=c++17"]
  }]
}
