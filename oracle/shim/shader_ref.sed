# oracle/shim/shader_ref.sed -- TEST INFRASTRUCTURE ONLY.
# The only edits made (on the fly, nothing is written back) to /root/reference/VolumeRenderer.cs so that
# g++ accepts the GLSL text as the body of `struct Shader` (oracle/Makefile, _ref/libshader_ref.so):
# directives and qualifiers that have no C++ spelling.  No expression or statement is touched.
/^#version/d
/^layout *(local_size/d
s/^layout *([^)]*) *uniform Camera/struct Camera_block/
s/^layout *([^)]*) *uniform //
# forward declarations at file scope (a class member function cannot be declared twice)
/^[A-Za-z0-9_]\+ [A-Za-z0-9_]\+(.*);[[:space:]]*$/d
# `out T name` parameters are references
s/\bout \+\([A-Za-z_][A-Za-z0-9_]*\) \+/\1\& /g
