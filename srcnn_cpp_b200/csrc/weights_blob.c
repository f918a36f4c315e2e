/* weights_blob.c -- embeds srcnn_cpp_b200/data/srcnn_weights.bin into the shared library, the way
 * the reference compiles convdata.h into bin/srcnn.  SRCNN_WEIGHTS_BIN is set by the Makefile. */
#ifndef SRCNN_WEIGHTS_BIN
#error "define SRCNN_WEIGHTS_BIN to the path of srcnn_weights.bin"
#endif
__asm__(
    ".section .rodata\n"
    ".balign 64\n"
    ".global srcnn_weights_blob\n"
    ".type srcnn_weights_blob, @object\n"
    "srcnn_weights_blob:\n"
    ".incbin \"" SRCNN_WEIGHTS_BIN "\"\n"
    "srcnn_weights_blob_end:\n"
    ".size srcnn_weights_blob, .-srcnn_weights_blob\n"
    ".balign 4\n"
    ".global srcnn_weights_blob_size\n"
    ".type srcnn_weights_blob_size, @object\n"
    "srcnn_weights_blob_size:\n"
    ".int srcnn_weights_blob_end - srcnn_weights_blob\n"
    ".size srcnn_weights_blob_size, 4\n"
    ".section .note.GNU-stack,\"\",@progbits\n"
    ".text\n");
