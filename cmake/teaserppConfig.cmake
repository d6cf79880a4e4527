# teaserppConfig.cmake -- drop-in for the reference's CMake package (cmake/teaserppConfig.cmake there):
# `find_package(teaserpp REQUIRED)` yields the imported targets the reference's examples link,
#     teaserpp::teaser_registration   and   teaserpp::teaser_io
# (examples/teaser_cpp_ply/CMakeLists.txt:7,16), both backed by libpsulvsb_b200.so and this repo's include/.
# Use:  cmake -Dteaserpp_DIR=<this repo>/cmake ...   (build the library first: python __graft_entry__.py)
get_filename_component(_PSULVSB_ROOT "${CMAKE_CURRENT_LIST_DIR}/.." ABSOLUTE)
set(_PSULVSB_LIB
    "${_PSULVSB_ROOT}/probabilistic-self-update-line-vector-set-based-point-cloud-registration_b200/libpsulvsb_b200.so")
if(NOT EXISTS "${_PSULVSB_LIB}")
  message(FATAL_ERROR "teaserpp (psulvsb-b200): ${_PSULVSB_LIB} is missing -- run `python __graft_entry__.py` first")
endif()
foreach(_t teaser_registration teaser_io)
  if(NOT TARGET teaserpp::${_t})
    add_library(teaserpp::${_t} SHARED IMPORTED)
    set_target_properties(teaserpp::${_t} PROPERTIES
      IMPORTED_LOCATION "${_PSULVSB_LIB}"
      IMPORTED_NO_SONAME TRUE
      INTERFACE_INCLUDE_DIRECTORIES "${_PSULVSB_ROOT}/include")
  endif()
endforeach()
set(teaserpp_FOUND TRUE)
