/* Empty stand-in so the reference translation unit can be compiled without
 * libusb-1.0 development headers (oracle build only; test infrastructure). */
