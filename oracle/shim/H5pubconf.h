/* oracle/shim: nanopolish_eventalign.cpp includes the HDF5 build configuration header; nothing from it is used. */
